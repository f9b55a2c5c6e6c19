"""a synthetic track — camera poses on a helix around a point cloud; per frame
Arrsac(1e-5, init 16384, max_cand 1024, est/block 256) + LambdaTwist on the visible FeatureWorldMatches
(cv-sfm single-view registration, cv-sfm/src/lib.rs:1619-1631), then LinearEigenTriangulator on every
landmark with >= 3 observations (cv-geom/src/triangulation.rs:82-130).  Reduced size: GPU == CPU port exactly;
full size: ground-truth properties."""
import numpy as np
import pytest

import cv_b200
from oracle import pyoracle as O
from tests.geom_util import rot_angle, world_homog
from tests.synth import helix_track, landmark_observations as _landmark_observations

pytestmark = pytest.mark.gpu


def _register(frames, cloud, seed, init, cand, use_oracle):
    out = []
    for k, fr in enumerate(frames):
        world = world_homog(cloud[fr["ids"]])
        if use_oracle:
            cfg = O.arrsac_cfg(1e-5, initialization_hypotheses=init, max_candidate_hypotheses=cand, estimations_per_block=256)
            out.append(O.arrsac(cfg, 1, fr["bearing"], world, O.rng_xoshiro(seed + k)))
        else:
            ars = (cv_b200.Arrsac(1e-5, cv_b200.Xoshiro256PlusPlus(seed + k)).initialization_hypotheses(init)
                   .max_candidate_hypotheses(cand).estimations_per_block(256))
            out.append(ars.model_inliers(cv_b200.LambdaTwist(), fr["bearing"], world))
    return out


def test_track_reduced_matches_cpu_port():
    cloud, frames = helix_track(0, 6, 800, 400, 0.2)
    got = _register(frames, cloud, 0, 1024, 128, use_oracle=False)
    want = _register(frames, cloud, 0, 1024, 128, use_oracle=True)
    for g, w, fr in zip(got, want, frames):
        assert (g is None) == (w is None) and g is not None
        assert np.array_equal(g[2], w[2])
        assert np.allclose(g[0], w[0], atol=1e-8) and np.allclose(g[1], w[1], atol=1e-8)
        assert rot_angle(g[0], fr["R"]) < 1e-2 and np.linalg.norm(g[1] - fr["t"]) < 0.1     # minimal-sample (3-point) pose
    ids, poses, bearings, offsets = _landmark_observations(frames, got)
    assert len(ids) > 100
    pts, ok = cv_b200.LinearEigenTriangulator().triangulate_batch(poses, bearings, offsets)
    for n, l in enumerate(ids):
        w = O.triangulate_linear_eigen(poses[offsets[n]:offsets[n + 1]], bearings[offsets[n]:offsets[n + 1]])
        assert ok[n] == (w is not None)
        if w is not None:
            assert np.allclose(pts[n], w, atol=1e-9)


def test_track_full_size_properties():
    # config 5 sizes per frame (2k visible of 20k, 20 % outliers, 16384 / 1024 hypotheses); 16 of the 256 poses keep the test short
    cloud, frames = helix_track(1, 16, 20000, 2000, 0.2)
    regs = _register(frames, cloud, 100, 16384, 1024, use_oracle=False)
    for reg, fr in zip(regs, frames):
        assert reg is not None
        assert fr["good"][reg[2]].mean() > 0.98 and len(reg[2]) > 0.5 * fr["good"].sum()
        assert rot_angle(reg[0], fr["R"]) < 1e-2 and np.linalg.norm(reg[1] - fr["t"]) < 0.1
        res = cv_b200.residuals_world_to_camera([(reg[0], reg[1])], fr["bearing"], world_homog(cloud[fr["ids"]]))[0]
        assert np.array_equal(np.where(res < 1e-5)[0], reg[2])
    ids, poses, bearings, offsets = _landmark_observations(frames, regs)
    assert len(ids) > 500
    pts, ok = cv_b200.LinearEigenTriangulator().triangulate_batch(poses, bearings, offsets)
    # the reference's cheirality rule compares the world-frame ray with the direction origin -> point (not camera -> point,
    # triangulation.rs:120-127), so landmarks between the world origin and a camera are rejected: about half on this orbit
    assert 0.3 < ok.mean() < 0.8
    xyz = pts[ok, :3] / pts[ok, 3:4]
    err = np.linalg.norm(xyz - cloud[np.array(ids)[ok]], axis=1)
    assert np.median(err) < 0.05 and np.percentile(err, 95) < 0.4
