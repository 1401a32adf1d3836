"""Raw pinned host -> device copy bandwidth of the box (one cudaMemcpyAsync of 4 GiB, and 16 chunks of 256 MiB)."""
import time, torch
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, chunks in (("one copy", 1), ("16 chunks", 16), ("64 chunks", 64)):
    c = n // chunks
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        for k in range(chunks):
            d[k * c:(k + 1) * c].copy_(h[k * c:(k + 1) * c], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("%s: %.1f GB/s" % (name, n / dt / 1e9))
t0 = time.perf_counter(); h2 = d.cpu(); print("D2H pageable: %.1f GB/s" % (n / (time.perf_counter() - t0) / 1e9))
