"""bincode images of the records in the reference's `VSlamData` checkpoint, host side only.

vslam-sandbox saves and restores its map with `bincode::serialize_into(file, &vslam.data)` / `bincode::deserialize_from`
(vslam-sandbox/src/main.rs:80-89, 166-172).  That is bincode 1.x's default configuration: little-endian fixed-width integers,
`usize` as u64, sequence and map lengths as u64, `Option` as a u8 tag, fixed arrays / tuples / structs as their fields in order with
no framing.  This module writes and reads, in that encoding, every record of `VSlamData` whose layout the reference tree itself
defines (cv-sfm/src/lib.rs:59-203 and the cv-core / cv-pinhole / akaze types they hold):

  Feature {bearing, response, color}            cv-sfm/src/lib.rs:61-65            31 bytes
  akaze::KeyPoint                               akaze/src/lib.rs:71-93             36 bytes
  cv_core::KeyPoint(Point2<f64>)                cv-core/src/keypoint.rs:25         16 bytes
  WorldToCamera / CameraToWorld / CameraToCamera / WorldToWorld (IsometryMatrix3<f64>)   cv-core/src/pose.rs:182,210,237,304   96 bytes
  CameraIntrinsics, CameraIntrinsicsK1Distortion cv-pinhole/src/lib.rs:32-36,150-153   40 / 48 bytes
  slot-map keys (FeedKey, FrameKey, ViewKey, LandmarkKey, ...)                     8 bytes {idx: u32, version: u32}
  View, Landmark, Feed, ThreeViewConstraint, BundleAdjustment                      cv-sfm/src/lib.rs:100-153

nalgebra's serde forms (its `serde-serialize` feature, which cv-core turns on): a statically sized matrix is its column-major
elements as nested fixed arrays (no length), `Unit<V>` / `Point` / `Rotation` / `Translation` are transparent over their matrix,
`Isometry` is the struct {rotation, translation}.

NOT covered -- the containers: `DenseSlotMap` (slotmap), `HggLite` (hgg), `HammingHasher` (hamming-lsh) and `BitArray` (bitarray)
are external crates whose sources are not in the reference tree, so the byte image of a whole `VSlamData` file cannot be restated
from the tree; the record images here are what those containers hold.  No golden checkpoint exists upstream: parity unpinned,
the tests check the encoding rules, sizes and round trips.
"""
import struct

import numpy as np

# packed numpy views of the fixed-size records (vectorised encode / decode of what the hot path produces)
FEATURE_DTYPE = np.dtype([("bearing", "<f8", (3,)), ("response", "<f4"), ("color", "u1", (3,))])                   # 31 bytes
AKAZE_KEYPOINT_DTYPE = np.dtype([("point", "<f4", (2,)), ("response", "<f4"), ("size", "<f4"), ("octave", "<u8"),
                                 ("class_id", "<u8"), ("angle", "<f4")])                                               # 36 bytes
POSE_DTYPE = np.dtype([("rotation", "<f8", (3, 3)), ("translation", "<f8", (3,))])       # rotation stored column by column, 96 bytes
assert FEATURE_DTYPE.itemsize == 31 and AKAZE_KEYPOINT_DTYPE.itemsize == 36 and POSE_DTYPE.itemsize == 96


class Writer:
    """bincode 1.x default-configuration encoder."""

    def __init__(self):
        self.parts = []

    def raw(self, b):
        self.parts.append(bytes(b))
        return self

    def u8(self, v): return self.raw(struct.pack("<B", v))
    def u32(self, v): return self.raw(struct.pack("<I", v))
    def u64(self, v): return self.raw(struct.pack("<Q", v))
    usize = u64
    def f32(self, v): return self.raw(struct.pack("<f", v))
    def f64(self, v): return self.raw(struct.pack("<d", v))
    def boolean(self, v): return self.u8(1 if v else 0)

    def option(self, v, put):
        if v is None:
            return self.u8(0)
        self.u8(1)
        put(self, v)
        return self

    def seq(self, items, put):
        """Vec<T> / slice: u64 length, then the elements."""
        items = list(items)
        self.u64(len(items))
        for it in items:
            put(self, it)
        return self

    def mapping(self, items, put_key, put_value):
        """HashMap<K, V>: u64 length, then (key, value) in iteration order (a HashMap's order is unspecified; any order decodes)."""
        items = list(items.items()) if hasattr(items, "items") else list(items)
        self.u64(len(items))
        for k, v in items:
            put_key(self, k)
            put_value(self, v)
        return self

    def bytes(self):
        return b"".join(self.parts)


class Reader:
    """bincode 1.x default-configuration decoder; raises ValueError on truncation or an invalid tag."""

    def __init__(self, data):
        self.d = memoryview(bytes(data))
        self.p = 0

    def raw(self, n):
        if self.p + n > len(self.d):
            raise ValueError("bincode: unexpected end of input")
        b = self.d[self.p:self.p + n]
        self.p += n
        return bytes(b)

    def u8(self): return struct.unpack("<B", self.raw(1))[0]
    def u32(self): return struct.unpack("<I", self.raw(4))[0]
    def u64(self): return struct.unpack("<Q", self.raw(8))[0]
    usize = u64
    def f32(self): return struct.unpack("<f", self.raw(4))[0]
    def f64(self): return struct.unpack("<d", self.raw(8))[0]

    def boolean(self):
        t = self.u8()
        if t > 1:
            raise ValueError("bincode: invalid bool")
        return t == 1

    def option(self, get):
        t = self.u8()
        if t > 1:
            raise ValueError("bincode: invalid Option tag")
        return get(self) if t else None

    def seq(self, get):
        n = self.u64()
        if n > len(self.d) - self.p:              # every element takes at least one byte in these records
            raise ValueError("bincode: sequence length exceeds input")
        return [get(self) for _ in range(n)]

    def mapping(self, get_key, get_value):
        n = self.u64()
        if n > len(self.d) - self.p:
            raise ValueError("bincode: map length exceeds input")
        out = {}
        for _ in range(n):
            k = get_key(self)
            out[k] = get_value(self)
        return out

    def done(self):
        return self.p == len(self.d)


# ---- records ---------------------------------------------------------------------------------------------------------------
def put_key(w, key):
    """slot-map key (FeedKey, FrameKey, ViewKey, LandmarkKey, ReconstructionKey, ConstraintKey): {idx: u32, version: u32}."""
    idx, version = key
    w.u32(idx).u32(version)


def get_key(r):
    return (r.u32(), r.u32())


def put_pose(w, rotation, translation):
    """IsometryMatrix3<f64> (the newtype poses of cv-core/src/pose.rs): rotation column by column, then the translation."""
    R = np.asarray(rotation, np.float64).reshape(3, 3)
    t = np.asarray(translation, np.float64).reshape(3)
    w.raw(np.ascontiguousarray(R.T, "<f8").tobytes()).raw(t.astype("<f8").tobytes())


def get_pose(r):
    v = np.frombuffer(r.raw(96), "<f8")
    return v[:9].reshape(3, 3).T.copy(), v[9:].copy()


def put_intrinsics(w, focals, principal_point, skew=0.0):
    """cv_pinhole::CameraIntrinsics {focals, principal_point, skew} (cv-pinhole/src/lib.rs:32-36)."""
    w.f64(focals[0]).f64(focals[1]).f64(principal_point[0]).f64(principal_point[1]).f64(skew)


def get_intrinsics(r):
    fx, fy, cx, cy, skew = (r.f64() for _ in range(5))
    return {"focals": (fx, fy), "principal_point": (cx, cy), "skew": skew}


def put_intrinsics_k1(w, focals, principal_point, skew=0.0, k1=0.0):
    """cv_pinhole::CameraIntrinsicsK1Distortion {simple_intrinsics, k1} (cv-pinhole/src/lib.rs:150-153)."""
    put_intrinsics(w, focals, principal_point, skew)
    w.f64(k1)


def get_intrinsics_k1(r):
    d = get_intrinsics(r)
    d["k1"] = r.f64()
    return d


def put_view(w, frame, rotation, translation, landmarks):
    """View {frame: FrameKey, pose: WorldToCamera, landmarks: Vec<LandmarkKey>} (cv-sfm/src/lib.rs:111-118)."""
    put_key(w, frame)
    put_pose(w, rotation, translation)
    w.seq(landmarks, put_key)


def get_view(r):
    frame = get_key(r)
    R, t = get_pose(r)
    return {"frame": frame, "rotation": R, "translation": t, "landmarks": r.seq(get_key)}


def put_landmark(w, observations):
    """Landmark {observations: HashMap<ViewKey, usize>} (cv-sfm/src/lib.rs:103-106)."""
    w.mapping(observations, put_key, lambda w_, v: w_.usize(v))


def get_landmark(r):
    return r.mapping(get_key, lambda r_: r_.usize())


def put_feed(w, intrinsics_k1, frames):
    """Feed {intrinsics: CameraIntrinsicsK1Distortion, frames: Vec<FrameKey>} (cv-sfm/src/lib.rs:123-128)."""
    put_intrinsics_k1(w, **intrinsics_k1)
    w.seq(frames, put_key)


def get_feed(r):
    return {"intrinsics": get_intrinsics_k1(r), "frames": r.seq(get_key)}


def put_three_view_constraint(w, views, poses):
    """ThreeViewConstraint {views: [ViewKey; 3], poses: [IsometryMatrix3<f64>; 2]} (cv-sfm/src/lib.rs:157-162): arrays carry no length."""
    if len(views) != 3 or len(poses) != 2:
        raise ValueError("three views and two poses expected")
    for v in views:
        put_key(w, v)
    for R, t in poses:
        put_pose(w, R, t)


def get_three_view_constraint(r):
    return {"views": [get_key(r) for _ in range(3)], "poses": [get_pose(r) for _ in range(2)]}


def put_bundle_adjustment(w, reconstruction, updated_views, removed_views):
    """BundleAdjustment {reconstruction, updated_views: Vec<(ViewKey, WorldToCamera)>, removed_views: Vec<ViewKey>} (cv-sfm/src/lib.rs:145-152)."""
    put_key(w, reconstruction)
    w.seq(updated_views, lambda w_, e: (put_key(w_, e[0]), put_pose(w_, e[1][0], e[1][1])))
    w.seq(removed_views, put_key)


def get_bundle_adjustment(r):
    return {"reconstruction": get_key(r), "updated_views": r.seq(lambda r_: (get_key(r_), get_pose(r_))), "removed_views": r.seq(get_key)}


# ---- vectorised forms of what the hot path produces --------------------------------------------------------------------------
def features_to_bytes(bearings, responses, colors):
    """n `Feature` records (31 bytes each, cv-sfm/src/lib.rs:61-65) back to back: what `Frame::descriptor_features` holds per
    descriptor.  bearings: (n, 3) unit vectors from `CameraIntrinsics.calibrate(...).bearing()`, responses: the keypoints'
    detector responses, colors: (n, 3) u8 sampled at the keypoints (cv-sfm/src/lib.rs:599-640)."""
    b = np.asarray(bearings, np.float64).reshape(-1, 3)
    rec = np.zeros(len(b), FEATURE_DTYPE)
    rec["bearing"] = b
    rec["response"] = np.asarray(responses, np.float32).reshape(len(b))
    rec["color"] = np.asarray(colors, np.uint8).reshape(len(b), 3)
    return rec.tobytes()


def features_from_bytes(data):
    if len(data) % FEATURE_DTYPE.itemsize:
        raise ValueError("not a whole number of Feature records")
    return np.frombuffer(data, FEATURE_DTYPE).copy()


def akaze_keypoints_to_bytes(keypoints):
    """`Vec<akaze::KeyPoint>` (akaze/src/lib.rs:71-93) from the library's KP_DTYPE records: u64 length, then 36 bytes per keypoint
    (`octave` and `class_id` are `usize`, so 8 bytes each in bincode)."""
    k = np.asarray(keypoints)
    rec = np.zeros(len(k), AKAZE_KEYPOINT_DTYPE)
    rec["point"][:, 0] = k["x"]; rec["point"][:, 1] = k["y"]
    rec["response"] = k["response"]; rec["size"] = k["size"]; rec["octave"] = k["octave"]; rec["class_id"] = k["class_id"]; rec["angle"] = k["angle"]
    return struct.pack("<Q", len(k)) + rec.tobytes()


def akaze_keypoints_from_bytes(data):
    from ._lib import KP_DTYPE
    r = Reader(data)
    n = r.u64()
    rec = np.frombuffer(r.raw(n * AKAZE_KEYPOINT_DTYPE.itemsize), AKAZE_KEYPOINT_DTYPE)
    k = np.zeros(n, KP_DTYPE)
    k["x"] = rec["point"][:, 0]; k["y"] = rec["point"][:, 1]
    for f in ("response", "size", "octave", "class_id", "angle"):
        k[f] = rec[f]
    return k


def poses_to_bytes(rotations, translations):
    """n IsometryMatrix3<f64> records back to back (96 bytes each); rotations (n, 3, 3) row-major in, column-major on the wire."""
    R = np.asarray(rotations, np.float64).reshape(-1, 3, 3)
    rec = np.zeros(len(R), POSE_DTYPE)
    rec["rotation"] = R.transpose(0, 2, 1)
    rec["translation"] = np.asarray(translations, np.float64).reshape(len(R), 3)
    return rec.tobytes()


def poses_from_bytes(data):
    rec = np.frombuffer(data, POSE_DTYPE)
    return rec["rotation"].transpose(0, 2, 1).copy(), rec["translation"].copy()
