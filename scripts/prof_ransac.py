"""Per-kernel breakdown of the two-view ARRSAC call on the bench's frame pair (vslam-sandbox configuration)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv_b200
from tests.synth import synth_frame, warp_frame

a = synth_frame(0); b = warp_frame(a, 1000)
ak = cv_b200.Akaze(maximum_features=5000)
k1, d1 = ak.extract_from_gray_float_image(a)
k2, d2 = ak.extract_from_gray_float_image(b)
pairs = np.array(cv_b200.symmetric_matching(d1, d2, 24), dtype=np.int64).reshape(-1, 2)
K = cv_b200.CameraIntrinsics(focals=(1000.0, 1000.0), principal_point=(960.0, 540.0))
ba = K.calibrate_keypoints(k1[pairs[:, 0]]); bb = K.calibrate_keypoints(k2[pairs[:, 1]])
ctx = ak.ctx
def run():
    ars = cv_b200.Arrsac(1e-7, cv_b200.Xoshiro256PlusPlus(0), ctx=ctx).initialization_hypotheses(8192).max_candidate_hypotheses(1024)
    t0 = time.perf_counter(); r = ars.model_inliers(cv_b200.EightPoint(), ba, bb); return r, (time.perf_counter() - t0) * 1e3
run()
r, ms = run()
ctx.profile(True)
r2, ms2 = run()
rep = ctx.profile_report()
ctx.profile(False)
print(json.dumps({"matches": len(pairs), "inliers": len(r[2]), "ms": ms, "ms_profiled": ms2, "kernels": rep}, indent=1))
