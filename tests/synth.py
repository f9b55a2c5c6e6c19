"""Seeded synthetic inputs: textured frames with blobs/corners, a warped second
view with ground-truth correspondences, and random 486-bit descriptors."""
import numpy as np


def _value_noise(rng, h, w, cell):
    gh, gw = h // cell + 3, w // cell + 3
    g = rng.random((gh, gw), dtype=np.float32)
    ys = (np.arange(h, dtype=np.float32) / cell)
    xs = (np.arange(w, dtype=np.float32) / cell)
    y0 = ys.astype(np.int32); x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy); fx = fx * fx * (3 - 2 * fx)
    a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def synth_frame(seed, h=1080, w=1920, nblobs=9000):
    """Band-limited value noise (4 octaves) plus random Gaussian blobs and boxes; float32 in [0,1]."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float32)
    for cell, amp in ((64, 0.30), (24, 0.20), (8, 0.12), (3, 0.05)):
        img += amp * _value_noise(rng, h, w, cell)
    n = nblobs
    cx = rng.uniform(8, w - 8, n); cy = rng.uniform(8, h - 8, n)
    sg = rng.uniform(1.2, 6.0, n); am = rng.uniform(-0.45, 0.45, n)
    for i in range(n):
        r = int(3 * sg[i]) + 1
        x0, x1 = max(int(cx[i]) - r, 0), min(int(cx[i]) + r + 1, w)
        y0, y1 = max(int(cy[i]) - r, 0), min(int(cy[i]) + r + 1, h)
        yy = np.arange(y0, y1, dtype=np.float32)[:, None] - np.float32(cy[i])
        xx = np.arange(x0, x1, dtype=np.float32)[None, :] - np.float32(cx[i])
        if i % 3 == 0:   # box corner-ish structure
            img[y0:y1, x0:x1] += np.float32(am[i]) * ((np.abs(yy) < sg[i]) & (np.abs(xx) < sg[i]))
        else:
            img[y0:y1, x0:x1] += np.float32(am[i]) * np.exp(-(yy * yy + xx * xx) / np.float32(2 * sg[i] * sg[i]))
    img -= img.min()
    img /= max(float(img.max()), 1e-6)
    return np.ascontiguousarray(img, dtype=np.float32)


def warp_frame(img, seed, shift=(3.4, -2.2), noise=1.0 / 255):
    """Second view: sub-pixel translation + slight scale about the centre (bilinear) + N(0, noise)."""
    rng = np.random.default_rng(seed)
    h, w = img.shape
    s = 1.01
    ys = (np.arange(h, dtype=np.float32)[:, None] - h / 2) / s + h / 2 - np.float32(shift[1])
    xs = (np.arange(w, dtype=np.float32)[None, :] - w / 2) / s + w / 2 - np.float32(shift[0])
    ys = np.clip(ys, 0, h - 1.001); xs = np.clip(xs, 0, w - 1.001)
    y0 = ys.astype(np.int32); x0 = xs.astype(np.int32)
    fy = ys - y0; fx = xs - x0
    out = (img[y0, x0] * (1 - fx) + img[y0, x0 + 1] * fx) * (1 - fy) + (img[y0 + 1, x0] * (1 - fx) + img[y0 + 1, x0 + 1] * fx) * fy
    out = out + rng.normal(0, noise, size=out.shape).astype(np.float32)
    return np.ascontiguousarray(np.clip(out, 0, 1), dtype=np.float32)


def random_descriptors(n, seed):
    """Uniform random 486-bit descriptors in BitArray<64> layout (top 26 bits zero)."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    d[:, 61:] = 0
    d[:, 60] &= 0x3F
    return d


# ---- a synthetic track (camera on a helix around a point cloud)
def helix_track(seed, n_poses, n_points, visible, outlier_frac, noise_px=0.3, focal=1000.0):
    from tests.geom_util import rot_from_euler as _rot_from_euler, unit
    rng = np.random.default_rng(seed)
    cloud = rng.uniform(-3, 3, (n_points, 3)) + np.array([0, 0, 0.0])
    frames = []
    for f in range(n_poses):
        ang = 2 * np.pi * f / 64.0
        centre = np.array([9 * np.cos(ang), 9 * np.sin(ang), -2 + 4.0 * f / max(n_poses - 1, 1)])
        fwd = unit(-centre); up = np.array([0, 0, 1.0]); right = unit(np.cross(fwd, up)); dwn = np.cross(fwd, right)
        R = np.stack([right, dwn, fwd]) @ _rot_from_euler(*rng.uniform(-0.02, 0.02, 3))       # world -> camera rotation
        t = -R @ centre
        cam = cloud @ R.T + t
        vis = np.where((cam[:, 2] > 1.0) & (np.abs(cam[:, 0] / cam[:, 2]) < 0.9) & (np.abs(cam[:, 1] / cam[:, 2]) < 0.5))[0]
        vis = rng.choice(vis, min(visible, len(vis)), replace=False)
        px = cam[vis, :2] / cam[vis, 2:3] + rng.normal(0, noise_px / focal, (len(vis), 2))
        bearing = unit(np.concatenate([px, np.ones((len(vis), 1))], 1))
        good = np.ones(len(vis), bool)
        bad = rng.choice(len(vis), int(outlier_frac * len(vis)), replace=False)
        bearing[bad] = unit(np.concatenate([rng.uniform(-0.9, 0.9, (len(bad), 1)), rng.uniform(-0.5, 0.5, (len(bad), 1)),
                                            np.ones((len(bad), 1))], 1))
        good[bad] = False
        frames.append(dict(R=R, t=t, ids=vis, bearing=bearing, good=good))
    return cloud, frames


def landmark_observations(frames, regs, min_obs=3, max_obs=None):
    """Observations (pose, bearing) of every landmark seen as an inlier in >= min_obs registered frames (at most the first
    `max_obs` of them: a feature track of bounded length)."""
    obs = {}
    for fr, reg in zip(frames, regs):
        if reg is None:
            continue
        for i in reg[2]:
            lst = obs.setdefault(int(fr["ids"][i]), [])
            if max_obs is None or len(lst) < max_obs:
                lst.append(((reg[0], reg[1]), fr["bearing"][i]))
    ids = sorted(l for l, v in obs.items() if len(v) >= min_obs)
    poses, bearings, offsets = [], [], [0]
    for l in ids:
        for p, b in obs[l]:
            poses.append(p); bearings.append(b)
        offsets.append(len(poses))
    return ids, poses, np.array(bearings), offsets
