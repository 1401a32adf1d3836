"""First-contact GPU script: hooks vs oracle, smoke, a mid-size batch with stage timings."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import spherical_sfm_b200 as S
import oracle as O
import __graft_entry__ as G

def Em(m): return np.array([m[0], m[1], m[2], m[1], -m[0], m[3], m[4], m[5], 0.0])
def step(name, fn):
    t = time.time()
    try:
        fn(); print("[ok] %s (%.2fs)" % (name, time.time() - t), flush=True)
    except Exception:
        print("[FAIL] %s" % name); traceback.print_exc(); sys.stdout.flush()

eng = S.Engine(0)
orc = O.load()
thr2 = (2 / 600) ** 2
print("fp32 peak TFLOP/s", eng.measure_fp32_peak(), flush=True)

def t_shuffle():
    sizes = np.array([450, 21, 3, 1000, 2, 1, 7], np.int32); tg = np.minimum(sizes, 21)
    assert (eng.lo_shuffle(42, sizes, tg) == orc.lo_shuffle(42, sizes, tg)).all()
step("lo_shuffle", t_shuffle)

def t_solve():
    rng = S.problems.make_rng(5, 0)
    pr = S.problems.make_problem(rng, 500, False, None, 1 / 600, 100, 20)
    samples = np.array([S.sample(0, 0, i, 3, 500) for i in range(256)], np.int32)
    for kind in (0, 1, 2):
        models, nm = eng.minimal_solve(pr.rays, samples, kind)
        worst = 0
        for s in range(len(samples)):
            nmo, mo = orc.solve(pr.rays, samples[s], kind)
            assert nmo == nm[s], (kind, s, nmo, nm[s])
            for a in models[s][:nm[s]]:
                if np.isnan(a).any(): continue
                d = min(min(np.linalg.norm(a - b), np.linalg.norm(a + b)) for b in mo[:nmo])
                worst = max(worst, d)
        print("   kind", kind, "worst model diff vs oracle", worst)
        assert worst < 1e-5
step("minimal_solve", t_solve)

def t_score():
    rng = S.problems.make_rng(6, 0)
    pr = S.problems.make_problem(rng, 3000, False, None, 1 / 600, 1500, 20)
    samples = np.array([S.sample(0, 0, i, 3, 3000) for i in range(300)], np.int32)
    models, nm = eng.minimal_solve(pr.rays, samples, 0)
    m6 = models.reshape(-1, 6)
    s32, c32, ms = eng.score(m6, pr.rays, thr2)
    so, co, _ = orc.score_batch(m6, pr.rays, thr2)
    rel = np.abs(s32 - so) / so
    print("   f32 score rel err max %.3e median %.3e; count mismatches %d of %d (max abs %d); kernel %.3f ms" % (np.nanmax(rel), np.nanmedian(rel), (c32 != co).sum(), len(co), np.abs(c32 - co).max(), ms))
    E9 = np.array([Em(m) for m in m6])
    se, ce = eng.score_exact(E9, pr.rays, thr2)
    ok = ~np.isnan(so)
    print("   f64 exact: count equal", (ce[ok] == co[ok]).all(), "score rel", np.max(np.abs(se[ok] - so[ok]) / so[ok]))
    assert (ce[ok] == co[ok]).all()
step("score", t_score)

def t_lsq_decomp():
    rng = S.problems.make_rng(8, 0)
    pr = S.problems.make_problem(rng, 400, False, None, 1 / 600, 100, 20)
    inl = np.nonzero(pr.inlier_mask)[0].astype(np.int32)
    E0 = pr.E / np.linalg.norm(pr.E)
    r, t = orc.decompose(E0)
    Es = [orc.make_E(r + 0.01 * rng.standard_normal(3)).reshape(9) for _ in range(8)]
    samples = [inl[: 21 + 10 * i] for i in range(8)]
    out = eng.least_squares(pr.rays, samples, np.array(Es))
    worst = 0
    for i in range(8):
        Eo, it, term, costs = orc.lm_refit(pr.rays, samples[i], Es[i])
        worst = max(worst, np.abs(Eo.reshape(9) - out[i]).max())
    print("   LM worst diff", worst)
    rr, tt = eng.decompose(np.array(Es))
    wd = max(np.abs(orc.decompose(Es[i])[0] - rr[i]).max() for i in range(8))
    print("   decompose worst diff", wd)
    assert worst < 1e-8 and wd < 1e-10
step("least_squares/decompose", t_lsq_decomp)

step("smoke", G.smoke)

def t_batch(P, N, outl, opt, name):
    rays, offsets, probs = S.problems.make_batch(99, P, N, noise=1 / 600, outlier_frac=outl, max_angle_deg=20.0)
    t0 = time.time(); res, flags = eng.estimate_pairs(rays, offsets, opt); wall = time.time() - t0
    st = eng.stats()
    mism = 0
    nchk = min(P, 24)
    oopt = O.default_options()
    for f, _ in opt._fields_:
        of = {"solver": "solver_kind", "fixed_budget": "legacy_budget", "fixed_prob_success": "legacy_prob_success"}.get(f, f)
        if hasattr(oopt, of): setattr(oopt, of, getattr(opt, f))
    for p in range(nchk):
        ref, inl = orc.estimate_pair(rays[offsets[p]:offsets[p + 1]], oopt, p)
        fl = np.zeros(N, np.uint8); fl[inl] = 1
        same = (int(res["num_iterations"][p]) == ref.num_iterations and int(res["best_num_inliers"][p]) == ref.best_num_inliers and int(res["number_lo_iterations"][p]) == ref.number_lo_iterations and (flags[offsets[p]:offsets[p + 1]] == fl).all())
        if not same:
            mism += 1
            print("   mismatch pair", p, res["num_iterations"][p], ref.num_iterations, res["best_num_inliers"][p], ref.best_num_inliers, res["number_lo_iterations"][p], ref.number_lo_iterations)
    ev = int(res["evals"].sum())
    print("   %s: P=%d N=%d wall %.3fs device %.2f ms (solve %.2f score %.2f chain %.2f) rounds %d launches %d; useful evals %.3e -> %.3e evals/s (device), executed %.3e exact %.3e; iters mean %.1f; oracle mismatches %d/%d" % (
        name, P, N, wall, st.total_ms, st.solve_ms, st.score_ms, st.chain_ms, st.rounds, st.kernel_launches, ev, ev / (st.total_ms * 1e-3), st.evals_executed, st.evals_exact, res["num_iterations"].mean(), mism, nchk), flush=True)
    assert mism == 0

step("batch C1-like", lambda: t_batch(64, 1000, 0.5, S.pipeline_options(thr2), "pipeline"))
step("batch C3-like", lambda: t_batch(2048, 1500, 0.7, S.pipeline_options(thr2), "pipeline 70%"))
step("batch default LO", lambda: t_batch(64, 600, 0.5, S.default_options(squared_inlier_threshold=thr2), "default LO"))
step("batch vanilla", lambda: t_batch(64, 600, 0.5, S.default_options(squared_inlier_threshold=thr2, driver=1), "vanilla"))
step("batch legacy/fast", lambda: t_batch(256, 2000, 0.3, S.default_options(squared_inlier_threshold=thr2, driver=2, solver=2, fixed_budget=512), "legacy fast"))
step("batch ragged/tiny", lambda: None)
eng.close()
