"""Host-side logic above the C ABI that needs no GPU: the registration matching rules with the k-NN injected (CPU)."""
import numpy as np
import pytest

from cv_b200.sfm_match import landmark_matches
from oracle import pyoracle as O


def _scene(seed, n_land=300, n_views=3, per_view=180, n_new=200):
    rng = np.random.default_rng(seed)
    proto = rng.integers(0, 256, (n_land, 64), dtype=np.uint8)
    proto[:, 60] &= 0x3F; proto[:, 61:] = 0

    def noisy(d, flips):
        d = d.copy()
        for r in range(len(d)):
            for bit in rng.choice(486, flips, replace=False):
                d[r, bit >> 3] ^= 1 << (bit & 7)
        return d
    views, lv, oc = [], {l: set() for l in range(n_land + 60)}, {l: 0 for l in range(n_land + 60)}
    for v in range(n_views):
        ids = rng.choice(n_land, per_view, replace=False).copy()
        desc = noisy(proto[ids], 12)
        if v > 0:
            ids[:15] = n_land + np.arange(15) + 15 * (v - 1)
        views.append((desc, ids))
        for l in ids:
            lv[int(l)].add(v); oc[int(l)] += 1
    return noisy(proto[rng.choice(n_land, n_new, replace=False)], 10), views, lv, oc


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_landmark_matches_rules_with_injected_knn(seed):
    new, views, lv, oc = _scene(seed)
    got = landmark_matches(new, views, 24, lv, oc, knn=O.hamming_knn)
    want = O.landmark_matches_ref(new, views, 24, lv, oc)
    assert got == want and len(got) > 50
    # every kept landmark is claimed exactly once, merge pairs never share a view, order is by observation count
    claimed = [l for m in got for l in m[0]]
    assert len(claimed) == len(set(claimed))
    assert all(not (lv[m[0][0]] & lv[m[0][1]]) for m in got if len(m[0]) == 2)
    counts = [sum(oc[l] for l in m[0]) for m in got]
    assert counts == sorted(counts, reverse=True)
    # a stricter margin can only remove unique matches
    strict = landmark_matches(new, views, 60, lv, oc, knn=O.hamming_knn)
    assert sum(len(m[0]) == 1 for m in strict) <= sum(len(m[0]) == 1 for m in got)


def test_landmark_matches_edge_cases():
    new, views, lv, oc = _scene(3)
    assert landmark_matches(new[:0], views, knn=O.hamming_knn) == []
    assert landmark_matches(new, [], knn=O.hamming_knn) == []
    with pytest.raises(ValueError):
        landmark_matches(new, [(views[0][0], views[0][1][:-1])], knn=O.hamming_knn)
    with pytest.raises(ValueError):           # fewer than three distinct candidate landmarks: the reference unwraps (panics)
        landmark_matches(new[:5], [(views[0][0][:2], views[0][1][:2])], knn=O.hamming_knn)


def test_pack_layout_uneven_shards():
    """config 4 with num_frames % world != 0: padded frames have count 0 and occupy no rows of the packed all-gather."""
    from cv_b200.multi import pack_layout
    from cv_b200 import dist as D
    world, F = 4, 10
    per = -(-F // world)
    counts = np.zeros((world, per), np.int64)
    for g in range(F):
        counts[g % world, g // world] = 100 + g
    off, maxtot = pack_layout(counts)
    assert maxtot == int(counts.sum(1).max()) and off.shape == (world, per + 1)
    for r in range(world):
        assert off[r, 0] == 0 and (np.diff(off[r]) == counts[r]).all()
    assert counts[2, 2] == 0 and counts[3, 2] == 0              # ranks 2, 3 hold two frames: their third slot is padding
    assert sorted(sum((D.shard_frames(F, r, world) for r in range(world)), [])) == list(range(F))
