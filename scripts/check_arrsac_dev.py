"""Device-resident ARRSAC vs the round-1 host-driven driver (CVB_ARRSAC_HOST=1) and timing, on the bench's frame pair
and on synthetic two-view / P3P scenes."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv_b200
from tests.geom_util import two_view_scene

def run(est, a, b, thr, seed, host, **kw):
    os.environ["CVB_ARRSAC_HOST"] = "1" if host else "0"
    rng = cv_b200.Xoshiro256PlusPlus(seed)
    ars = cv_b200.Arrsac(thr, rng)
    for k, v in kw.items():
        getattr(ars, k)(v)
    t0 = time.perf_counter()
    r = ars.model_inliers(est, a, b)
    ms = (time.perf_counter() - t0) * 1e3
    return r, ms, [int(x) for x in rng.state.s]

def compare(name, est, a, b, thr, seed, **kw):
    rd, msd, sd = run(est, a, b, thr, seed, False, **kw)
    rd, msd, sd = run(est, a, b, thr, seed, False, **kw)
    rh, msh, sh = run(est, a, b, thr, seed, True, **kw)
    ok = (rd is None) == (rh is None)
    if rd is not None and rh is not None:
        ok = ok and np.array_equal(rd[2], rh[2]) and np.array_equal(rd[0], rh[0]) and np.array_equal(rd[1], rh[1])
    ok = ok and sd == sh
    print(json.dumps({"case": name, "n": len(a), "equal": bool(ok), "rng_equal": sd == sh, "inliers_dev": None if rd is None else len(rd[2]),
                      "inliers_host": None if rh is None else len(rh[2]), "ms_dev": round(msd, 3), "ms_host": round(msh, 3)}), flush=True)
    return ok

allok = True
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bench_pair0.npz")) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bench_pair0.npz")) else None
rng = np.random.default_rng(5)
for n, of, noise, thr, kw in [(11, 0.0, 0.0, 0.1, {}), (300, 0.3, 1e-4, 1e-6, {}), (1000, 0.3, 5e-5, 1e-7, dict(initialization_hypotheses=512, max_candidate_hypotheses=128)),
                              (1500, 0.3, 5e-5, 1e-7, dict(initialization_hypotheses=8192, max_candidate_hypotheses=1024)),
                              (3500, 0.1, 2e-5, 1e-7, dict(initialization_hypotheses=8192, max_candidate_hypotheses=1024)),
                              (777, 0.5, 1e-4, 1e-6, dict(block_size=50, initialization_blocks=3, estimations_per_block=32))]:
    R, t, a, b, good = two_view_scene(rng, n, outlier_frac=of, noise=noise)
    allok &= compare(f"two_view_{n}", cv_b200.EightPoint(), a, b, thr, 0, **kw)
if z is not None:
    allok &= compare("bench_pair0", cv_b200.EightPoint(), z["ba"], z["bb"], 1e-7, 0, initialization_hypotheses=8192, max_candidate_hypotheses=1024)
    ctx = cv_b200._lib.default_context(0)
    ctx.profile(True)
    os.environ["CVB_ARRSAC_HOST"] = "0"
    run(cv_b200.EightPoint(), z["ba"], z["bb"], 1e-7, 0, False, initialization_hypotheses=8192, max_candidate_hypotheses=1024)
    print(json.dumps(ctx.profile_report()))
    ctx.profile(False)
# P3P (single-view configuration of vslam-sandbox)
from tests.geom_util import pnp_scene
R, t, bear, world, good = pnp_scene(rng, 2000, outlier_frac=0.2, noise=1e-4)
allok &= compare("p3p_2000", cv_b200.LambdaTwist(), bear, world, 1e-5, 0, initialization_hypotheses=16384, max_candidate_hypotheses=1024, estimations_per_block=256)
print("ALL_EQUAL", allok)
sys.exit(0 if allok else 1)
