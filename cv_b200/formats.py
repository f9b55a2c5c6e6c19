"""On-disk formats either side of the hot path, host side only.

  write_akaze_dump / read_akaze_dump <- akaze/examples/akaze.rs:10-33: `<stem>_kps.csv` ("x, y, angle, size, octave, class_id" with
        Rust's `{}` float formatting) and `<stem>_descs.txt` (64 bytes as `{:08b}` joined by `_`): the format the reference's own
        golden vectors would be dumped in
  export_ply                         <- cv-sfm/src/export.rs:20-136: ASCII PLY, vertices (double x y z, uchar red green blue), cameras
        as a magenta centre + 4 image-plane corners (+ 4 triangles when camera_faces)

The `ply-rs` crate is not in the reference tree; header and line layout follow the PLY specification as that crate documents it
(parity of the text unpinned)."""
import numpy as np

from ._lib import KP_DTYPE

CAMERA_COLOR = (255, 0, 255)


def rust_float(x):
    """Rust `format!("{}", x)` for f32 / f64: shortest round-trip digits, positional (never an exponent), no trailing `.0`."""
    x = x if isinstance(x, (np.floating,)) else np.float64(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    return np.format_float_positional(x, unique=True, trim="-")


def write_akaze_dump(stem, keypoints, descriptors):
    """Writes `<stem>_kps.csv` and `<stem>_descs.txt` (akaze/examples/akaze.rs:13-31); returns the two paths."""
    kps = np.asarray(keypoints)
    desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 64)
    if len(kps) != len(desc):
        raise ValueError("one descriptor per keypoint expected")
    kp_path, desc_path = f"{stem}_kps.csv", f"{stem}_descs.txt"
    with open(kp_path, "w") as f:
        for k in kps:
            f.write(f"{rust_float(k['x'])}, {rust_float(k['y'])}, {rust_float(k['angle'])}, {rust_float(k['size'])}, "
                    f"{int(k['octave'])}, {int(k['class_id'])}\n")
    with open(desc_path, "w") as f:
        for d in desc:
            f.write("_".join(format(int(b), "08b") for b in d) + "\n")
    return kp_path, desc_path


def read_akaze_dump(stem):
    """Inverse of write_akaze_dump -> (keypoints KP_DTYPE [response = 0: the dump does not carry it], descriptors [N, 64] uint8)."""
    rows = [line.split(",") for line in open(f"{stem}_kps.csv").read().splitlines() if line.strip()]
    kps = np.zeros(len(rows), KP_DTYPE)
    for i, r in enumerate(rows):
        if len(r) != 6:
            raise ValueError(f"line {i + 1}: expected 6 fields")
        kps[i]["x"], kps[i]["y"], kps[i]["angle"], kps[i]["size"] = (np.float32(v) for v in r[:4])
        kps[i]["octave"], kps[i]["class_id"] = int(r[4]), int(r[5])
    lines = [line for line in open(f"{stem}_descs.txt").read().splitlines() if line.strip()]
    desc = np.zeros((len(lines), 64), np.uint8)
    for i, line in enumerate(lines):
        parts = line.split("_")
        if len(parts) != 64 or any(len(p) != 8 for p in parts):
            raise ValueError(f"descriptor line {i + 1}: expected 64 groups of 8 bits")
        desc[i] = [int(p, 2) for p in parts]
    if len(desc) != len(kps):
        raise ValueError("keypoint / descriptor files disagree in length")
    return kps, desc


def export_ply(writer, points_and_colors, cameras=(), camera_faces=False):
    """cv-sfm/src/export.rs:20-136.  points_and_colors: [(xyz, (r, g, b))]; cameras: dicts / objects with optical_center,
    up_direction, forward_direction, focal_length.  Vertices: per camera the centre and the corners (up,right), (up,-right),
    (-up,-right), (-up,right), then the points; faces (if requested): four triangles per camera."""
    vertices, faces = [], []

    def add_vertex(p, c):
        vertices.append((np.asarray(p, np.float64), tuple(int(v) for v in c)))
        return len(vertices) - 1

    for cam in cameras:
        get = cam.get if isinstance(cam, dict) else lambda k, cam=cam: getattr(cam, k)
        oc, up, fw = (np.asarray(get(k), np.float64) for k in ("optical_center", "up_direction", "forward_direction"))
        fl = float(get("focal_length"))
        right = np.cross(fw, up)
        centre = add_vertex(oc, CAMERA_COLOR)
        up_right, up_left, down_left, down_right = [add_vertex(oc + fw * fl + float(u) * up * fl + float(r) * right * fl, CAMERA_COLOR)
                                                    for u, r in ((1, 1), (1, -1), (-1, -1), (-1, 1))]
        if camera_faces:
            faces += [(centre, down_right, up_right), (centre, up_right, up_left), (centre, up_left, down_left), (centre, down_left, down_right)]
    for p, c in points_and_colors:
        add_vertex(p, c)
    out = ["ply", "format ascii 1.0", "comment Exported from rust-cv/vslam-sandbox", f"element vertex {len(vertices)}",
           "property double x", "property double y", "property double z", "property uchar red", "property uchar green", "property uchar blue"]
    if camera_faces:
        out += [f"element face {len(faces)}", "property list uchar int vertex_index"]
    out.append("end_header")
    for p, c in vertices:
        out.append(f"{rust_float(p[0])} {rust_float(p[1])} {rust_float(p[2])} {c[0]} {c[1]} {c[2]}")
    if camera_faces:
        out += [f"3 {a} {b} {c}" for a, b, c in faces]
    text = "\n".join(out) + "\n"
    writer.write(text if "b" not in getattr(writer, "mode", "") else text.encode())
