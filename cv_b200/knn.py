"""Host-side mirror of space::LinearKnn + bitarray::Hamming and the match filters built on it."""
import ctypes as C

import numpy as np

from ._lib import default_context


def _desc(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim != 2 or a.shape[1] != 64:
        raise ValueError("descriptors must be [N, 64] uint8 (BitArray<64>)")
    return a


def hamming_knn(queries, database, k=2, ctx=None):
    """All-queries form of `LinearKnn{metric: Hamming, iter: database}.knn(q, k)`.

    Returns (index[N,k], distance[N,k]) uint32, ascending distance, ties -> lower index first
    (space 0.17 LinearKnn; call sites akaze/tests/estimate_pose.rs:82-86).  Missing neighbours
    (database smaller than k) are 0xffffffff.
    """
    q, db = _desc(queries), _desc(database)
    ctx = ctx or default_context(0)
    idx = np.empty((len(q), k), np.uint32)
    dist = np.empty((len(q), k), np.uint32)
    ctx.check(ctx.lib.cvb_hamming_knn(ctx.handle, q.ctypes.data, len(q), db.ctypes.data, len(db), k, idx.ctypes.data,
                                      dist.ctypes.data))
    return idx, dist


class LinearKnn:
    """space::LinearKnn { metric: Hamming, iter } -- `knn(query, num)` yields Neighbor{index, distance}."""

    def __init__(self, database, ctx=None):
        self.database = _desc(database)
        self.ctx = ctx

    def knn(self, query, num):
        idx, dist = hamming_knn(np.asarray(query, np.uint8).reshape(1, 64), self.database, num, self.ctx)
        return [(int(i), int(d)) for i, d in zip(idx[0], dist[0]) if i != 0xFFFFFFFF]

    def knn_batch(self, queries, num):
        return hamming_knn(queries, self.database, num, self.ctx)

    def nn(self, query):
        r = self.knn(query, 1)
        return r[0] if r else None


def lowe_ratio_matches(ds1, ds2, ratio=0.5, ctx=None):
    """akaze/tests/estimate_pose.rs:78-97 `match_descriptors`: 2-NN + Lowe ratio in f32, -> [(ix1, ix2), ...]."""
    idx, dist = hamming_knn(ds1, ds2, 2, ctx)
    ok = dist[:, 0].astype(np.float32) < dist[:, 1].astype(np.float32) * np.float32(ratio)
    return [(int(i), int(idx[i, 0])) for i in np.where(ok)[0]]


def matching(a, b, better_by=24, strict=False, ctx=None):
    """cv-sfm `matching` (cv-sfm/src/lib.rs:3097-3114: d0 + better_by <= d1); strict=True gives the
    tutorial rule d0 + 24 < d1 (chapter4 main.rs:99).  Returns an int64 array, -1 where None."""
    a, b = _desc(a), _desc(b)
    if len(a) < 2 or len(b) < 2:
        return np.zeros(0, np.int64)
    idx, dist = hamming_knn(a, b, 2, ctx)
    d0, d1 = dist[:, 0].astype(np.int64), dist[:, 1].astype(np.int64)
    good = (d0 + better_by < d1) if strict else (d0 + better_by <= d1)
    return np.where(good, idx[:, 0].astype(np.int64), -1)


def symmetric_matching(a, b, better_by=24, ctx=None):
    """cv-sfm `symmetric_matching` (cv-sfm/src/lib.rs:3116-3133): [[aix, bix], ...] in ascending aix."""
    a, b = _desc(a), _desc(b)
    ctx = ctx or default_context(0)
    cap = max(len(a), 1)
    pairs = np.empty((cap, 2), np.uint32)
    n = C.c_uint32()
    ctx.check(ctx.lib.cvb_match_symmetric(ctx.handle, a.ctypes.data, len(a), b.ctypes.data, len(b), better_by,
                                          pairs.ctypes.data, cap, C.byref(n)))
    return pairs[:n.value].astype(np.int64)


class HammingHasher:
    """hamming_lsh::HammingHasher<64, H>::new_with_codewords(codewords) (cv-sfm/src/lib.rs:205,216): `hash_bag(features)` sets, for
    every feature, the bit of its nearest codeword (first minimum on ties).  codewords: [H * 8, 64] uint8."""

    def __init__(self, codewords, ctx=None):
        self.codewords = _desc(codewords)
        if len(self.codewords) == 0 or len(self.codewords) % 32:
            raise ValueError("the number of codewords must be a positive multiple of 32")
        self.ctx = ctx

    def hash_bag(self, features):
        f = np.ascontiguousarray(features, np.uint8).reshape(-1, 64)
        ctx = self.ctx or default_context(0)
        out = np.zeros(len(self.codewords) // 8, np.uint8)
        ctx.check(ctx.lib.cvb_hash_bag(ctx.handle, f.ctypes.data, len(f), self.codewords.ctypes.data, len(self.codewords), out.ctypes.data))
        return out
