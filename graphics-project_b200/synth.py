"""Procedural synthetic frames for benchmarks and parity tests (data generation, not hot path).

The reference renders its test images with a moderngl/pygame simulator (run_simulation.py,
containers.py:42-78 random cube scene, shaders/default_flat.frag:46-61 white outline band).  That
renderer is out of scope; this module draws statistically similar frames on the
host with cv2's polygon rasteriser: unit cubes at integer grid positions, grey faces, near-white
outlines, simulator clear colour (run_simulation.py:88-90) as background.

Three scene kinds:
  "cubes" -- 12 cubes, free camera (BASELINE configs 2, 3)
  "voxel" -- 40-120 grid-aligned cubes (config 4)
  "cave"  -- voxel + 16x16 per-face block-texture noise (config 5)
Frames are deterministic: ``np.random.default_rng(seed)``, seed = frame index.
"""
import numpy as np
import cv2 as _cv  # drawing only (fillConvexPoly / polylines); never used for the detectors

BACKGROUND_BGR = (242, 153, 128)  # simulator clear colour (0.5, 0.6, 0.95) as BGR u8
OUTLINE_BGR = (235, 255, 255)

_CUBE_V = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], np.float64)
# faces as vertex indices (quads, consistent winding not required for the painter's algorithm)
_CUBE_F = np.array([[0, 1, 3, 2], [4, 5, 7, 6], [0, 1, 5, 4], [2, 3, 7, 6], [0, 2, 6, 4], [1, 3, 7, 5]])


def _camera(yaw, pitch, dist):
    """World->camera rotation R and translation t for a camera orbiting the origin."""
    eye = dist * np.array([np.cos(pitch) * np.cos(yaw), -np.sin(pitch), np.cos(pitch) * np.sin(yaw)])
    fwd = -eye / np.linalg.norm(eye)
    up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd])
    return R, -R @ eye


def make_frame(seed, height=1080, width=1920, kind="cubes"):
    """One BGR uint8 (height, width, 3) frame."""
    rng = np.random.default_rng(seed)
    img = np.empty((height, width, 3), np.uint8)
    img[:] = BACKGROUND_BGR
    if kind == "cubes":
        ncubes = 12
    else:
        ncubes = int(rng.integers(40, 121))
    pos = np.stack([rng.integers(-4, 5, ncubes), rng.integers(-2, 2, ncubes), rng.integers(-4, 5, ncubes)], 1)
    pos = np.unique(pos, axis=0)
    yaw = rng.uniform(0, 2 * np.pi)
    pitch = rng.uniform(-0.6, 0.2)
    R, t = _camera(yaw, pitch, 8.0 if kind == "cubes" else 11.0)
    f = height / 2.0
    cx, cy = width / 2.0, height / 2.0
    thick = max(1, int(round(3 * height / 400)))
    faces = []
    for p in pos:
        vc = (R @ (_CUBE_V + p - 0.5).T).T + t
        if np.any(vc[:, 2] < 0.5):
            continue
        pix = np.stack([f * vc[:, 0] / vc[:, 2] + cx, f * vc[:, 1] / vc[:, 2] + cy], 1)
        for fi in _CUBE_F:
            faces.append((vc[fi, 2].mean(), pix[fi], int(rng.integers(25, 70))))
    faces.sort(key=lambda a: -a[0])
    for _, quad, grey in faces:
        q = np.round(quad * 16).astype(np.int32)
        if kind == "cave":
            # block texture: fill, then 4x4 sub-quads with +-40 grey noise (cheap 16x16-like texture)
            n = 4
            a, b, c, d = quad
            for i in range(n):
                for j in range(n):
                    u0, u1, v0, v1 = i / n, (i + 1) / n, j / n, (j + 1) / n

                    def P(u, v):
                        return (a * (1 - u) + b * u) * (1 - v) + (d * (1 - u) + c * u) * v

                    sub = np.round(np.array([P(u0, v0), P(u1, v0), P(u1, v1), P(u0, v1)]) * 16).astype(np.int32)
                    g = int(np.clip(grey + rng.integers(-40, 41), 0, 255))
                    _cv.fillConvexPoly(img, sub, (g, g, g), lineType=_cv.LINE_AA, shift=4)
        else:
            _cv.fillConvexPoly(img, q, (grey, grey, grey), lineType=_cv.LINE_AA, shift=4)
        _cv.polylines(img, [q], True, OUTLINE_BGR, thickness=thick, lineType=_cv.LINE_AA, shift=4)
    return img


def make_batch(n, height=1080, width=1920, kind="cubes", first_seed=0, unique=None):
    """(n, height, width, 3) uint8 batch; ``unique`` < n tiles that many distinct frames cyclically."""
    unique = n if unique is None else min(unique, n)
    base = [make_frame(first_seed + i, height, width, kind) for i in range(unique)]
    out = np.empty((n, height, width, 3), np.uint8)
    for i in range(n):
        out[i] = base[i % unique]
    return out
