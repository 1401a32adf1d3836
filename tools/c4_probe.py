"""Config C4 timing probe: P pairs x 1000 correspondences, 50 % outliers, six-point shared focal, VanillaMSAC."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, sixpt_oracle as X, spherical_sfm_b200 as S

if __name__ == "__main__":
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rays, offsets, f, R, t = S.problems.make_sixpt_batch(1, P, 1000)
    eng = S.Engine(0)
    lo = len(sys.argv) > 2 and sys.argv[2] == "lo"  # LO-MSAC with RansacLib's default LO schedule instead of VanillaMSAC
    opt = S.default_options(squared_inlier_threshold=4.0, driver=S.DRIVER_LO_MSAC if lo else S.DRIVER_VANILLA_MSAC, solver=S.SOLVER_SIXPT_FOCAL,
                            sixpt_focal_scoring=1, random_seed=1234)
    eng.upload(rays, offsets)
    for rep in range(3):
        t0 = time.time(); eng.run(opt); dt = time.time() - t0
        st = eng.stats()
        print(f"rep {rep}: {dt*1e3:.1f} ms  solve {st.solve_ms:.1f} score {st.score_ms:.1f} chain {st.chain_ms:.1f} rounds {st.rounds} "
              f"launches {st.kernel_launches} fp32 evals {st.evals_executed:.3e} exact {st.evals_exact:.3e}")
    res, flags = eng.download()
    its = res["num_iterations"]; print("iterations mean", its.mean(), "max", its.max(), "useful evals", res["evals"].sum(),
          "pairs/s", P / dt, "evals/s", res["evals"].sum() / dt)
    ok = (np.abs(res["focal"] / f - 1) < 0.5).mean(); print("focal within 50%:", ok, "within 5%:", (np.abs(res["focal"] / f - 1) < 0.05).mean(), "inlier ratio mean", res["inlier_ratio"].mean(),
          "LO calls per pair", res["number_lo_iterations"].mean())
