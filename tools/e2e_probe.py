import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import spherical_sfm_b200 as S, bench
P, N = 124750, 1500
rays_dev, offsets, _ = bench.make_batch_torch(P, N, 0.7, 1234, "cuda")
rays_host = torch.empty(rays_dev.shape, dtype=torch.float64, pin_memory=True); rays_host.copy_(rays_dev); del rays_dev; torch.cuda.empty_cache()
rays = rays_host.numpy()
eng = S.Engine(0); opt = S.pipeline_options(bench.THR2)
def T(f, n=3):
    f(); ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return min(ts) * 1e3
print("upload (blocking) ms", T(lambda: eng.upload(rays, offsets)))
print("run ms", T(lambda: eng.run(opt)))
print("download(flags) ms", T(lambda: eng.download(True)))
print("download(no flags) ms", T(lambda: eng.download(False)))
rt = torch.empty(P * S.RESULT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True); ft = torch.empty(P * N, dtype=torch.uint8, pin_memory=True)
r_ = rt.numpy().view(S.RESULT_DTYPE); f_ = ft.numpy()
print("estimate_pairs pipelined pinned-out ms", T(lambda: eng.estimate_pairs(rays, offsets, opt, out_results=r_, out_flags=f_)))
print("estimate_pairs pipelined ms", T(lambda: eng.estimate_pairs(rays, offsets, opt)))
print("estimate_pairs pipelined, no flags ms", T(lambda: eng.estimate_pairs(rays, offsets, opt, want_flags=False)))
os.environ["SSFM_NO_PIPELINE"] = "1"
print("estimate_pairs plain ms", T(lambda: eng.estimate_pairs(rays, offsets, opt)))
st = eng.stats(); print("stats", st.total_ms, st.solve_ms, st.score_ms, st.chain_ms, st.rounds, st.kernel_launches)
