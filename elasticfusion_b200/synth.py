"""Deterministic synthetic RGB-D data (ICL-NUIM-shaped) for parity tests and bench.py.

The reference ships no sample data and the box has no network, so every workload in
BASELINE.json's `configs` is generated here: an analytic pinhole render of a textured axis-aligned box room
(5 x 3 x 4 m) seen from a smooth Lissajous trajectory, depth as uint16 millimetres (the input format of
ElasticFusion::processFrame, Core/ElasticFusion.h:58-75), colour as uint8 RGB.

Conventions: camera x right, y down, z forward; pixel (u, v) samples the ray ((u-cx)/fx, (v-cy)/fy, 1)
(integer pixel coordinates, as Core/Cuda/cudafuncs.cu:138-139 back-projects). Poses are 4x4 camera-to-world.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

ROOM_MIN = np.array([-2.5, -1.5, -2.0])
ROOM_MAX = np.array([2.5, 1.5, 2.0])


@dataclass(frozen=True)
class Intrinsics:
    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float

    def scaled(self, s: int) -> "Intrinsics":
        return Intrinsics(self.width * s, self.height * s, self.fx * s, self.fy * s, self.cx * s, self.cy * s)


# the reference app's default camera (MainController.cpp:37-42) and the ICL-NUIM one
K_DEFAULT = Intrinsics(640, 480, 528.0, 528.0, 320.0, 240.0)
K_ICLNUIM = Intrinsics(640, 480, 481.2, 480.0, 319.5, 239.5)


def rot_xyz(rx: float, ry: float, rz: float) -> np.ndarray:
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def pose(R: np.ndarray, t) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def trajectory(n: int, seed: int = 42, speed: float = 1.0) -> np.ndarray:
    """(n,4,4) camera-to-room poses: Lissajous sway, <= ~1 cm and <= ~0.5 deg per frame at speed 1."""
    rng = np.random.RandomState(seed)
    ph = rng.uniform(0, 2 * np.pi, size=6)
    k = np.arange(n) * speed
    w = 2 * np.pi / 300.0
    px = 0.3 + 0.25 * (np.sin(w * k + ph[0]) - np.sin(ph[0]))
    py = 0.3 + 0.08 * (np.sin(2 * w * k + ph[1]) - np.sin(ph[1]))
    pz = -0.3 + 0.20 * (np.sin(1.5 * w * k + ph[2]) - np.sin(ph[2]))
    yaw = np.deg2rad(35.0) + np.deg2rad(9.0) * (np.sin(w * k + ph[3]) - np.sin(ph[3]))
    pitch = np.deg2rad(-14.0) + np.deg2rad(4.0) * (np.sin(2 * w * k + ph[4]) - np.sin(ph[4]))
    roll = np.deg2rad(2.0) * (np.sin(1.5 * w * k + ph[5]) - np.sin(ph[5]))
    out = np.empty((n, 4, 4))
    for i in range(n):
        # yaw about y (down), pitch about x; camera looks along +z
        out[i] = pose(rot_xyz(0, yaw[i], 0) @ rot_xyz(-pitch[i], 0, 0) @ rot_xyz(0, 0, roll[i]), [px[i], py[i], pz[i]])
    return out


def corner_trajectory(n: int, seed: int = 42, speed: float = 1.0) -> np.ndarray:
    """(n,4,4) camera-to-room poses looking DOWN into a floor corner of the room from ~1.5 m: three mutually orthogonal planes
    are in view (the floor on a few per cent of the pixels at the extremes of the sway), inside the 3 m depth cut-off over the
    first ~100 / speed frames, so the geometric term alone constrains all six degrees of freedom (the default trajectory sees two walls beyond the cut-off for most of its pixels: point-to-plane ICP slides along
    them and only the photometric term holds the pose). Same Lissajous sway, <= ~1 cm and <= ~0.5 deg per frame at speed 1."""
    rng = np.random.RandomState(seed)
    ph = rng.uniform(0, 2 * np.pi, size=6)
    k = np.arange(n) * speed
    w = 2 * np.pi / 300.0
    px = 1.1 + 0.22 * (np.sin(w * k + ph[0]) - np.sin(ph[0]))
    py = 0.45 + 0.08 * (np.sin(2 * w * k + ph[1]) - np.sin(ph[1]))
    pz = 0.6 + 0.18 * (np.sin(1.5 * w * k + ph[2]) - np.sin(ph[2]))
    yaw = np.deg2rad(45.0) + np.deg2rad(9.0) * (np.sin(w * k + ph[3]) - np.sin(ph[3]))
    pitch = np.deg2rad(32.0) + np.deg2rad(4.0) * (np.sin(2 * w * k + ph[4]) - np.sin(ph[4]))
    roll = np.deg2rad(2.0) * (np.sin(1.5 * w * k + ph[5]) - np.sin(ph[5]))
    out = np.empty((n, 4, 4))
    for i in range(n):
        out[i] = pose(rot_xyz(0, yaw[i], 0) @ rot_xyz(pitch[i], 0, 0) @ rot_xyz(0, 0, roll[i]), [px[i], py[i], pz[i]])
    return out


def corner_sequence(n: int, K: Intrinsics = K_DEFAULT, seed: int = 42, noise: bool = True, speed: float = 1.0):
    """Like sequence(), over corner_trajectory()."""
    traj = corner_trajectory(n, seed=seed, speed=speed)
    T0inv = np.linalg.inv(traj[0])
    for i in range(n):
        rgb, depth, _, _ = render(traj[i], K, noise_seed=(seed * 100003 + i) if noise else None)
        yield rgb, depth, T0inv @ traj[i]


def _hash01(ix, iy, iz, salt):
    h = (ix.astype(np.int64) * 73856093) ^ (iy.astype(np.int64) * 19349663) ^ (iz.astype(np.int64) * 83492791) ^ salt
    h = (h ^ (h >> 13)) * 1274126177
    h = h ^ (h >> 16)
    return (h & 0xFFFF).astype(np.float64) / 65535.0


def texture(p: np.ndarray, seed: int = 1234) -> np.ndarray:
    """uint8 RGB (never 0) as a function of the room-frame hit point p[..., 3]: smooth shading + 0.3 m
    high-contrast tiles (the photometric term only uses strong edges: minimumGradientMagnitudes {5,3,1},
    Core/Utils/RGBDOdometry.cpp:112-114) + voxel-hashed grain."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    tile = (np.floor(x / 0.3) + np.floor(y / 0.3) + np.floor(z / 0.3)).astype(np.int64) & 1
    tile2 = (np.floor((x + 0.11) / 0.17) + np.floor((y + 0.07) / 0.23) + 2 * np.floor((z - 0.05) / 0.19)).astype(np.int64) & 1
    out = np.empty(p.shape[:-1] + (3,), dtype=np.float64)
    ph = np.random.RandomState(seed).uniform(0, 2 * np.pi, size=(3, 3))
    vx = np.floor(x / 0.01)
    vy = np.floor(y / 0.01)
    vz = np.floor(z / 0.01)
    for c in range(3):
        smooth = 30 * np.sin(3.1 * x + 2.3 * y + ph[c, 0]) + 22 * np.sin(4.7 * z - 1.9 * x + ph[c, 1]) + 14 * np.sin(7.3 * y + 5.1 * z + ph[c, 2])
        grain = 8.0 * (_hash01(vx, vy, vz, seed + 17 * c) - 0.5)
        out[..., c] = 128 + smooth + 38 * (tile - 0.5) * 2 * (0.8 + 0.2 * c) * 0.5 + 24 * (tile2 - 0.5) + grain
    return np.clip(np.rint(out), 1, 255).astype(np.uint8)


def render(T_room_c: np.ndarray, K: Intrinsics, noise_seed: int | None = None, max_mm: int = 20000,
           pixel_offset: float = 0.0):
    """Returns rgb (H,W,3) u8, depth (H,W) u16 mm, z (H,W) float64 exact, normal (H,W,3) camera-frame."""
    H, W = K.height, K.width
    u = np.arange(W, dtype=np.float64) + pixel_offset
    v = np.arange(H, dtype=np.float64) + pixel_offset
    uu, vv = np.meshgrid(u, v)
    d_c = np.stack([(uu - K.cx) / K.fx, (vv - K.cy) / K.fy, np.ones_like(uu)], axis=-1)
    R, o = T_room_c[:3, :3], T_room_c[:3, 3]
    d = d_c @ R.T
    best_t = np.full((H, W), np.inf)
    best_n = np.zeros((H, W, 3))
    with np.errstate(divide="ignore", invalid="ignore"):
        for axis in range(3):
            for bound, sign in ((ROOM_MIN[axis], 1.0), (ROOM_MAX[axis], -1.0)):
                t = (bound - o[axis]) / d[..., axis]
                ok = (t > 1e-6) & (t < best_t)
                hit = o + t[..., None] * d
                for a2 in range(3):
                    if a2 != axis:
                        ok &= (hit[..., a2] >= ROOM_MIN[a2] - 1e-9) & (hit[..., a2] <= ROOM_MAX[a2] + 1e-9)
                best_t = np.where(ok, t, best_t)
                nrm = np.zeros(3)
                nrm[axis] = sign
                best_n = np.where(ok[..., None], nrm, best_n)
    hit = o + best_t[..., None] * d
    rgb = texture(hit)
    z = best_t.copy()  # camera-frame z == t because d_c.z == 1
    zn = z
    if noise_seed is not None:
        rng = np.random.RandomState(noise_seed)
        zn = z + rng.standard_normal(z.shape) * 0.0012 * z * z
    mm = np.rint(zn * 1000.0)
    mm = np.where(np.isfinite(mm) & (mm > 0) & (mm <= max_mm), mm, 0)
    n_c = best_n @ R  # room -> camera: R^T n
    return rgb, mm.astype(np.uint16), z, n_c


def sequence(n: int, K: Intrinsics = K_DEFAULT, seed: int = 42, noise: bool = True, speed: float = 1.0):
    """Yields (rgb, depth_mm, T_w_c) with T expressed relative to the first camera (world == frame 0)."""
    traj = trajectory(n, seed=seed, speed=speed)
    T0inv = np.linalg.inv(traj[0])
    for i in range(n):
        rgb, depth, _, _ = render(traj[i], K, noise_seed=(seed * 100003 + i) if noise else None)
        yield rgb, depth, T0inv @ traj[i]


def write_klg(path: str, frames, timestamps_us=None) -> None:
    """Raw-payload .klg (Tools/RawLogReader.cpp:22-109): int32 numFrames; per frame int64 timestamp,
    int32 depthSize, int32 imageSize, depth bytes, image bytes. depthSize == 2*N and imageSize == 3*N mark
    uncompressed payloads (RawLogReader.cpp:80-97)."""
    frames = list(frames)
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(frames)))
        for i, (rgb, depth) in enumerate(frames):
            ts = int(timestamps_us[i]) if timestamps_us is not None else i * 33333
            db, ib = depth.astype("<u2").tobytes(), rgb.astype(np.uint8).tobytes()
            f.write(struct.pack("<qii", ts, len(db), len(ib)))
            f.write(db)
            f.write(ib)


def read_klg(path: str, width: int, height: int):
    """Reader for the raw-payload .klg written above (same layout as the reference reader)."""
    n_px = width * height
    with open(path, "rb") as f:
        (n,) = struct.unpack("<i", f.read(4))
        for _ in range(n):
            ts, dsz, isz = struct.unpack("<qii", f.read(16))
            if dsz != 2 * n_px or isz != 3 * n_px:
                raise ValueError("compressed .klg payloads (zlib depth / JPEG colour) are not supported")
            depth = np.frombuffer(f.read(dsz), dtype="<u2").reshape(height, width)
            rgb = np.frombuffer(f.read(isz), dtype=np.uint8).reshape(height, width, 3)
            yield ts, rgb, depth


def ate_rmse(est: np.ndarray, gt: np.ndarray) -> float:
    """Absolute trajectory error (translation RMSE, no alignment: both start at identity)."""
    d = est[:, :3, 3] - gt[:, :3, 3]
    return float(np.sqrt((d * d).sum(axis=1).mean()))


def room_surfels(n_target: int, T_w_room: np.ndarray, view_depth: float = 1.5, focal: float = 528.0, conf: float = 20.0,
                 time: float = 1.0) -> np.ndarray:
    """About n_target stable surfels tiling the six walls of the room on a regular grid, in the reference's 12-float Vertex
    layout (Core/Shaders/Vertex.cpp:22-41): pos.xyz conf | colour 0 initTime lastTime | normal.xyz radius, expressed in the
    world frame T_w_room maps the room into. Used by bench.py to make a map of the size BASELINE's configs state (5 M / 20 M)
    resident before timing. Colours come from `texture`, normals point away from the room's interior (the convention of
    geometry.glsl's central-difference normals: away from the viewer), the radius is what surfels.glsl assigns to a
    fronto-parallel surface seen from `view_depth` with focal length `focal`, widened so neighbouring discs overlap. All
    init times are equal, so any order satisfies the map-order invariant (App. A-22)."""
    ext = ROOM_MAX - ROOM_MIN
    area = 2.0 * (ext[0] * ext[1] + ext[0] * ext[2] + ext[1] * ext[2])
    s = float(np.sqrt(area / max(n_target, 1)))
    radius = max(view_depth / focal * np.sqrt(2.0), 0.95 * s)
    parts = []
    for axis in range(3):
        a1, a2 = [a for a in range(3) if a != axis]
        n1, n2 = max(1, int(round(ext[a1] / s))), max(1, int(round(ext[a2] / s)))
        u = ROOM_MIN[a1] + (np.arange(n1, dtype=np.float64) + 0.5) * (ext[a1] / n1)
        v = ROOM_MIN[a2] + (np.arange(n2, dtype=np.float64) + 0.5) * (ext[a2] / n2)
        uu, vv = np.meshgrid(u, v, indexing="ij")
        for bound, sign in ((ROOM_MIN[axis], -1.0), (ROOM_MAX[axis], 1.0)):
            p = np.empty((n1 * n2, 3))
            p[:, axis] = bound
            p[:, a1] = uu.ravel()
            p[:, a2] = vv.ravel()
            nrm = np.zeros(3)
            nrm[axis] = sign
            out = np.empty((n1 * n2, 12), np.float32)
            for c0 in range(0, len(p), 1 << 20):  # texture() in chunks: bounded temporaries
                c1 = min(len(p), c0 + (1 << 20))
                rgb = texture(p[c0:c1]).astype(np.int64)
                out[c0:c1, 4] = ((rgb[:, 0] << 16) + (rgb[:, 1] << 8) + rgb[:, 2]).astype(np.float32)
            out[:, 0:3] = (p @ T_w_room[:3, :3].T + T_w_room[:3, 3]).astype(np.float32)
            out[:, 3] = conf
            out[:, 5] = 0.0
            out[:, 6] = time
            out[:, 7] = time
            out[:, 8:11] = (T_w_room[:3, :3] @ nrm).astype(np.float32)
            out[:, 11] = radius
            parts.append(out)
    return np.ascontiguousarray(np.concatenate(parts, axis=0))
