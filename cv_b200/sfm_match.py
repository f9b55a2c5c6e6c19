"""cv-sfm's frame-registration matching stage re-typed over the exact GPU k-NN.

  landmark_matches <- VSlam::register_frame_subset, matching part       cv-sfm/src/lib.rs:1468-1562
                      (the reference runs one approximate `HggLite::knn(descriptor, 3)` per feature and view:
                      cv-sfm/src/lib.rs:1476-1490; here every view costs ONE batched N x M launch of the exact matcher)

Where the reference's result depends on `HashMap` iteration order (equal distances among the best three landmarks),
this module breaks ties by (distance, landmark id)."""
import numpy as np

from .knn import hamming_knn


def landmark_matches(new_descriptors, views, better_by=24, landmark_views=None, landmark_observations=None, ctx=None, knn=None):
    """new_descriptors[N, 64]; views = [(descriptors[M_v, 64], landmarks[M_v] int), ...] for every matched view.

    Returns [(landmarks, feature)] with landmarks a 1- or 2-tuple:
      * the best landmark when `d0 + better_by <= d1` (uniquely good);
      * the best two when `d1 + better_by <= d2` and they share no view (`landmark_views`: landmark -> iterable of view
        keys; omitted = never sharing) -- a merge candidate;
    then drops every match whose landmark is claimed by more than one feature, and stable-sorts by the summed observation
    count of its landmarks, descending (`landmark_observations`: landmark -> count; omitted = keep order).
    `knn(queries, database, k) -> (idx, dist)` replaces the GPU matcher (host-logic tests only)."""
    q = np.ascontiguousarray(new_descriptors, np.uint8)
    if len(views) == 0 or len(q) == 0:
        return []
    lm, dist = [], []
    for desc, landmarks in views:
        landmarks = np.asarray(landmarks, np.int64)
        if len(landmarks) != len(desc):
            raise ValueError("one landmark per view feature expected")
        idx, d = knn(q, desc, 3) if knn else hamming_knn(q, desc, 3, ctx)      # one N x M launch per view
        missing = idx == 0xFFFFFFFF
        lm.append(np.where(missing, -1, landmarks[np.where(missing, 0, idx)]))
        dist.append(np.where(missing, 1 << 20, d).astype(np.int64))
    lm = np.concatenate(lm, 1); dist = np.concatenate(dist, 1)         # [N, 3V]
    # keep the best instance of every landmark: sort by (landmark, distance), first of each run
    key = lm * (1 << 21) + dist
    order = np.argsort(key, axis=1, kind="stable")
    lm_s = np.take_along_axis(lm, order, 1); d_s = np.take_along_axis(dist, order, 1)
    first = np.ones_like(lm_s, bool)
    first[:, 1:] = lm_s[:, 1:] != lm_s[:, :-1]
    first &= lm_s >= 0
    # best three by (distance, landmark)
    key2 = np.where(first, d_s * (1 << 42) + lm_s, np.iinfo(np.int64).max)
    o2 = np.argsort(key2, axis=1, kind="stable")[:, :3]
    if o2.shape[1] < 3 or np.any(np.take_along_axis(key2, o2, 1) == np.iinfo(np.int64).max):
        raise ValueError("every feature needs three distinct candidate landmarks (the reference unwraps them)")
    bl = np.take_along_axis(lm_s, o2, 1); bd = np.take_along_axis(d_s, o2, 1)
    unique = bd[:, 0] + better_by <= bd[:, 1]
    merge = ~unique & (bd[:, 1] + better_by <= bd[:, 2])
    out = []
    for f in np.where(unique | merge)[0]:
        if unique[f]:
            out.append(((int(bl[f, 0]),), int(f)))
        else:
            a, b = int(bl[f, 0]), int(bl[f, 1])
            if landmark_views is None or not (set(landmark_views[a]) & set(landmark_views[b])):
                out.append(((a, b), int(f)))
    counts = {}
    for landmarks, _ in out:
        for l in landmarks:
            counts[l] = counts.get(l, 0) + 1
    out = [m for m in out if all(counts[l] == 1 for l in m[0])]
    if landmark_observations is not None:
        out.sort(key=lambda m: -sum(landmark_observations[l] for l in m[0]))      # sorted() is stable like sort_by_key
    return out
