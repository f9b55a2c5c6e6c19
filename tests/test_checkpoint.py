"""bincode record images of the reference's VSlamData checkpoint (cv_b200/checkpoint.py): encoding rules, sizes, round trips."""
import struct

import numpy as np
import pytest

from cv_b200 import checkpoint as ck
from cv_b200._lib import KP_DTYPE


def test_primitives_follow_bincode_default_configuration():
    w = ck.Writer()
    w.u8(7).u32(0x01020304).usize(5).f32(1.5).f64(-2.0).boolean(True)
    w.option(None, lambda w_, v: w_.u32(v)).option(9, lambda w_, v: w_.u32(v))
    w.seq([1, 2, 3], lambda w_, v: w_.u8(v))
    want = (b"\x07" + b"\x04\x03\x02\x01" + b"\x05" + b"\x00" * 7 + struct.pack("<f", 1.5) + struct.pack("<d", -2.0) + b"\x01"
            + b"\x00" + b"\x01\x09\x00\x00\x00" + b"\x03" + b"\x00" * 7 + b"\x01\x02\x03")
    assert w.bytes() == want
    r = ck.Reader(want)
    assert (r.u8(), r.u32(), r.usize(), r.f32(), r.f64(), r.boolean()) == (7, 0x01020304, 5, 1.5, -2.0, True)
    assert r.option(lambda r_: r_.u32()) is None and r.option(lambda r_: r_.u32()) == 9
    assert r.seq(lambda r_: r_.u8()) == [1, 2, 3] and r.done()


def test_truncated_and_invalid_input_is_rejected():
    with pytest.raises(ValueError):
        ck.Reader(b"\x01\x02").u32()
    with pytest.raises(ValueError):
        ck.Reader(b"\x02").option(lambda r_: r_.u8())
    with pytest.raises(ValueError):
        ck.Reader(struct.pack("<Q", 1 << 40)).seq(lambda r_: r_.u8())
    with pytest.raises(ValueError):
        ck.features_from_bytes(b"\x00" * 30)


def test_feature_record_is_31_bytes_in_field_order():
    b = ck.features_to_bytes([[0.0, 0.6, 0.8]], [0.25], [[1, 2, 3]])
    assert b == struct.pack("<dddf", 0.0, 0.6, 0.8, 0.25) + b"\x01\x02\x03" and len(b) == 31
    rng = np.random.default_rng(0)
    v = rng.standard_normal((100, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    resp = rng.uniform(0, 1, 100).astype(np.float32)
    col = rng.integers(0, 256, (100, 3), dtype=np.uint8)
    back = ck.features_from_bytes(ck.features_to_bytes(v, resp, col))
    assert np.array_equal(back["bearing"], v) and np.array_equal(back["response"], resp) and np.array_equal(back["color"], col)


def test_akaze_keypoints_usize_fields_take_eight_bytes():
    k = np.zeros(2, KP_DTYPE)
    k["x"] = [1.5, 100.25]; k["y"] = [2.5, 7.0]; k["response"] = [0.01, 0.02]; k["size"] = [4.8, 9.6]; k["octave"] = [0, 2]
    k["class_id"] = [3, 11]; k["angle"] = [-1.0, 3.0]
    b = ck.akaze_keypoints_to_bytes(k)
    assert len(b) == 8 + 2 * 36 and b[:8] == struct.pack("<Q", 2)
    assert b[8:8 + 36] == struct.pack("<ffffQQf", 1.5, 2.5, np.float32(0.01), np.float32(4.8), 0, 3, -1.0)
    back = ck.akaze_keypoints_from_bytes(b)
    for f in ("x", "y", "response", "size", "octave", "class_id", "angle"):
        assert np.array_equal(back[f], k[f])


def test_pose_is_column_major_rotation_then_translation():
    R = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    t = np.array([1.0, 2.0, 3.0])
    w = ck.Writer()
    ck.put_pose(w, R, t)
    assert w.bytes() == struct.pack("<12d", 0.0, 1.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 2.0, 3.0)
    assert w.bytes() == ck.poses_to_bytes(R[None], t[None])
    R2, t2 = ck.get_pose(ck.Reader(w.bytes()))
    assert np.array_equal(R2, R) and np.array_equal(t2, t)
    Rs, ts = ck.poses_from_bytes(ck.poses_to_bytes(np.stack([R, R.T]), np.stack([t, -t])))
    assert np.array_equal(Rs[1], R.T) and np.array_equal(ts[1], -t)


def test_intrinsics_records():
    w = ck.Writer()
    ck.put_intrinsics_k1(w, (1000.0, 1000.0), (960.0, 540.0), 0.0, -0.25)
    assert w.bytes() == struct.pack("<6d", 1000.0, 1000.0, 960.0, 540.0, 0.0, -0.25)
    assert ck.get_intrinsics_k1(ck.Reader(w.bytes())) == {"focals": (1000.0, 1000.0), "principal_point": (960.0, 540.0), "skew": 0.0, "k1": -0.25}


def test_map_records_round_trip():
    R = np.eye(3); t = np.array([0.5, 0.0, -1.0])
    w = ck.Writer()
    ck.put_view(w, (3, 1), R, t, [(0, 1), (7, 3)])
    ck.put_landmark(w, {(1, 1): 17, (2, 5): 4})
    ck.put_feed(w, {"focals": (700.0, 710.0), "principal_point": (320.0, 240.0), "skew": 0.1, "k1": 0.01}, [(0, 1), (1, 1), (2, 1)])
    ck.put_three_view_constraint(w, [(0, 1), (1, 1), (2, 1)], [(R, t), (R, -t)])
    ck.put_bundle_adjustment(w, (0, 1), [((4, 1), (R, t))], [(9, 2)])
    data = w.bytes()
    assert len(data) == (8 + 96 + 8 + 16) + (8 + 2 * 16) + (48 + 8 + 24) + (24 + 192) + (8 + 8 + 104 + 8 + 8)
    r = ck.Reader(data)
    v = ck.get_view(r)
    assert v["frame"] == (3, 1) and v["landmarks"] == [(0, 1), (7, 3)] and np.array_equal(v["translation"], t)
    assert ck.get_landmark(r) == {(1, 1): 17, (2, 5): 4}
    f = ck.get_feed(r)
    assert f["intrinsics"]["k1"] == 0.01 and f["frames"] == [(0, 1), (1, 1), (2, 1)]
    c = ck.get_three_view_constraint(r)
    assert c["views"] == [(0, 1), (1, 1), (2, 1)] and np.array_equal(c["poses"][1][1], -t)
    b = ck.get_bundle_adjustment(r)
    assert b["reconstruction"] == (0, 1) and b["updated_views"][0][0] == (4, 1) and b["removed_views"] == [(9, 2)]
    assert r.done()
    with pytest.raises(ValueError):
        ck.put_three_view_constraint(ck.Writer(), [(0, 1)], [])
