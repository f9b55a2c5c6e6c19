"""nalgebra-equivalent helpers for the geometry tests (plain numpy, f64)."""
import numpy as np


def rot_from_euler(roll, pitch, yaw):
    """nalgebra Rotation3::from_euler_angles(roll, pitch, yaw) = Rz(yaw) Ry(pitch) Rx(roll)."""
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def rot_from_scaled_axis(v):
    """nalgebra Rotation3::new(axisangle)."""
    v = np.asarray(v, np.float64)
    a = np.linalg.norm(v)
    if a == 0:
        return np.eye(3)
    u = v / a
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)


def rot_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.arccos(np.clip(c, -1, 1)))


def skew(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])


def unit(v):
    v = np.asarray(v, np.float64)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def world_homog(p):
    """Projective::from_point: xyz normalised, w = 1/|p|."""
    p = np.asarray(p, np.float64)
    n = np.linalg.norm(p, axis=-1, keepdims=True)
    return np.concatenate([p / n, 1.0 / n], axis=-1)


def two_view_scene(rng, n, outlier_frac=0.0, noise=0.0):
    """Random relative pose + n matches (unit bearings a, b); first (1-outlier_frac) are inliers. Shuffled."""
    R = rot_from_scaled_axis(rng.uniform(-1, 1, 3) * 0.25)
    t = unit(rng.uniform(-1, 1, 3))
    P = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(3, 8, n)], 1)
    Q = P @ R.T + t
    a, b = unit(P), unit(Q)
    if noise:
        a = unit(a + rng.normal(0, noise, a.shape)); b = unit(b + rng.normal(0, noise, b.shape))
    nout = int(n * outlier_frac)
    good = np.ones(n, bool)
    if nout:
        idx = rng.choice(n, nout, replace=False)
        b[idx] = unit(np.stack([rng.uniform(-0.4, 0.4, nout), rng.uniform(-0.4, 0.4, nout), np.ones(nout)], 1))
        good[idx] = False
    perm = rng.permutation(n)
    return R, t, a[perm], b[perm], good[perm]


def pnp_scene(rng, n, outlier_frac=0.0, noise=0.0):
    R = rot_from_euler(*rng.uniform(-0.4, 0.4, 3))
    t = rng.uniform(-0.5, 0.5, 3)
    C = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(2, 9, n)], 1)   # camera-frame points
    Wp = (C - t) @ R                                                                          # world = R^T (c - t)
    bearing = unit(C)
    if noise:
        bearing = unit(bearing + rng.normal(0, noise, bearing.shape))
    nout = int(n * outlier_frac)
    good = np.ones(n, bool)
    if nout:
        idx = rng.choice(n, nout, replace=False)
        bearing[idx] = unit(np.stack([rng.uniform(-0.5, 0.5, nout), rng.uniform(-0.5, 0.5, nout), np.ones(nout)], 1))
        good[idx] = False
    perm = rng.permutation(n)
    return R, t, bearing[perm], world_homog(Wp)[perm], good[perm]


def three_view_scene(rng, n, noise=0.0):
    """Centre camera = identity; returns [(R1, t1), (R2, t2)] CameraToCamera (centre -> first / second) and obs[n, 3, 3]."""
    poses = []
    for sgn in (1.0, -1.0):
        R = rot_from_scaled_axis(rng.uniform(-1, 1, 3) * 0.15)
        t = np.array([sgn * rng.uniform(0.4, 0.8), rng.uniform(-0.2, 0.2), rng.uniform(-0.1, 0.1)])
        poses.append((R, t))
    X = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(3, 8, n)], 1)
    obs = np.stack([unit(X), unit(X @ poses[0][0].T + poses[0][1]), unit(X @ poses[1][0].T + poses[1][1])], 1)
    if noise:
        obs = unit(obs + rng.normal(0, noise, obs.shape))
    return poses, obs


def perturb_pose(rng, pose, rot=0.01, trans=0.02):
    R, t = pose
    return rot_from_scaled_axis(rng.normal(0, 1, 3) * rot) @ R, t + rng.normal(0, 1, 3) * trans
