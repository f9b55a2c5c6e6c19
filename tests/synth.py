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
