"""Host-side formats next to the hot path (cv_b200/formats.py): the reference's keypoint / descriptor dump and its PLY export."""
import io

import numpy as np

from cv_b200._lib import KP_DTYPE
from cv_b200.formats import export_ply, read_akaze_dump, rust_float, write_akaze_dump


def test_rust_display_of_floats():
    # Rust `{}`: shortest digits that round-trip, positional, integers without a fraction
    assert rust_float(np.float32(1.0)) == "1" and rust_float(np.float32(0.1)) == "0.1" and rust_float(np.float32(-2.5)) == "-2.5"
    assert rust_float(np.float32(1e-7)) == "0.0000001" and rust_float(np.float32(16777216.0)) == "16777216"
    assert rust_float(np.float64(0.1) + np.float64(0.2)) == "0.30000000000000004"
    assert rust_float(np.float32("nan")) == "NaN" and rust_float(np.float32("inf")) == "inf" and rust_float(-np.float32("inf")) == "-inf"
    rng = np.random.default_rng(0)
    for v in rng.standard_normal(200).astype(np.float32) * np.float32(1000):
        assert np.float32(rust_float(v)) == v


def test_akaze_dump_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    n = 57
    kps = np.zeros(n, KP_DTYPE)
    kps["x"] = rng.uniform(0, 1920, n); kps["y"] = rng.uniform(0, 1080, n); kps["angle"] = rng.uniform(-3.2, 3.2, n)
    kps["size"] = rng.uniform(2, 60, n); kps["octave"] = rng.integers(0, 4, n); kps["class_id"] = rng.integers(0, 16, n)
    kps["response"] = rng.uniform(0, 1, n)
    desc = rng.integers(0, 256, (n, 64), dtype=np.uint8)
    stem = str(tmp_path / "frame")
    kp_path, desc_path = write_akaze_dump(stem, kps, desc)
    first = open(kp_path).readline().rstrip("\n").split(", ")
    assert len(first) == 6 and np.float32(first[0]) == kps["x"][0] and first[4] == str(int(kps["octave"][0]))
    line = open(desc_path).readline().rstrip("\n")
    assert len(line) == 64 * 8 + 63 and line.split("_")[3] == format(int(desc[0, 3]), "08b")      # `{x:08b}` joined by `_`
    k2, d2 = read_akaze_dump(stem)
    assert np.array_equal(d2, desc)
    for f in ("x", "y", "angle", "size", "octave", "class_id"):
        assert np.array_equal(k2[f], kps[f]), f
    assert np.all(k2["response"] == 0)


def test_export_ply_layout():
    cams = [dict(optical_center=[0.0, 0.0, 0.0], up_direction=[0.0, -1.0, 0.0], forward_direction=[0.0, 0.0, 1.0], focal_length=0.5)]
    pts = [((1.0, 2.0, 3.5), (10, 20, 30)), ((-0.25, 0.0, 1e-3), (0, 0, 255))]
    buf = io.StringIO()
    export_ply(buf, pts, cams, camera_faces=True)
    lines = buf.getvalue().splitlines()
    assert lines[:4] == ["ply", "format ascii 1.0", "comment Exported from rust-cv/vslam-sandbox", "element vertex 7"]
    assert lines[4:10] == ["property double x", "property double y", "property double z", "property uchar red", "property uchar green",
                           "property uchar blue"]
    assert lines[10:13] == ["element face 4", "property list uchar int vertex_index", "end_header"]
    assert lines[13] == "0 0 0 255 0 255"                       # camera centre, magenta
    # right = forward x up = (0,0,1) x (0,-1,0) = (1,0,0); corner (up, right) = centre + f*(fw + up + right)
    assert lines[14] == "0.5 -0.5 0.5 255 0 255" and lines[15] == "-0.5 -0.5 0.5 255 0 255"
    assert lines[18] == "1 2 3.5 10 20 30" and lines[19] == "-0.25 0 0.001 0 0 255"
    assert lines[20:] == ["3 0 4 1", "3 0 1 2", "3 0 2 3", "3 0 3 4"]
    buf2 = io.StringIO()
    export_ply(buf2, pts)
    assert "element face" not in buf2.getvalue() and buf2.getvalue().count("\n") == 11 + 2
