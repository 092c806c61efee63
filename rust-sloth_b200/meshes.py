"""Synthetic meshes for the benchmark and the parity tests.

``icosphere(f)``: class-I geodesic sphere of frequency f -- 20*f*f triangles,
vertices computed in float64 (barycentric split of the 20 faces of the unit
icosahedron, pushed to radius 1) and rounded to float32, counter-clockwise seen
from outside, one of 20 fixed colours per base face, soup in face-major order.
Deterministic, no RNG.  f = 708 gives the 10,025,280-triangle workload.
"""
from __future__ import annotations

import numpy as np

_PHI = (1.0 + 5.0 ** 0.5) / 2.0
_ICO_V = np.array([
    [-1, _PHI, 0], [1, _PHI, 0], [-1, -_PHI, 0], [1, -_PHI, 0],
    [0, -1, _PHI], [0, 1, _PHI], [0, -1, -_PHI], [0, 1, -_PHI],
    [_PHI, 0, -1], [_PHI, 0, 1], [-_PHI, 0, -1], [-_PHI, 0, 1]], np.float64)
_ICO_V /= np.linalg.norm(_ICO_V, axis=1, keepdims=True)
_ICO_F = np.array([
    [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
    [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
    [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
    [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)

FACE_COLOURS = np.array([
    [230, 25, 75], [60, 180, 75], [255, 225, 25], [0, 130, 200], [245, 130, 48],
    [145, 30, 180], [70, 240, 240], [240, 50, 230], [210, 245, 60], [250, 190, 212],
    [0, 128, 128], [220, 190, 255], [170, 110, 40], [255, 250, 200], [128, 0, 0],
    [170, 255, 195], [128, 128, 0], [255, 215, 180], [0, 0, 128], [128, 128, 128]], np.uint8)


def _face_soup(A, B, C, f: int) -> np.ndarray:
    """All f*f sub-triangles of one base face as (f*f, 9) float32."""
    def P(i, j):  # barycentric lattice point (i along AB, j along AC)
        i = i.astype(np.float64)[..., None]
        j = j.astype(np.float64)[..., None]
        p = (A * (f - i - j) + B * i + C * j) / f
        return p / np.linalg.norm(p, axis=-1, keepdims=True)

    out = []
    for r in range(f):          # row r: j = r, i = 0 .. f-r-1
        i = np.arange(f - r)
        j = np.full_like(i, r)
        up = np.concatenate([P(i, j), P(i + 1, j), P(i, j + 1)], axis=-1)
        if f - r - 1 > 0:
            i2 = np.arange(f - r - 1)
            j2 = np.full_like(i2, r)
            dn = np.concatenate([P(i2 + 1, j2), P(i2 + 1, j2 + 1), P(i2, j2 + 1)], axis=-1)
            # interleave up/down so that neighbours in memory are neighbours on the sphere
            row = np.empty((2 * (f - r) - 1, 9), np.float64)
            row[0::2] = up
            row[1::2] = dn
        else:
            row = up
        out.append(row)
    return np.concatenate(out).astype(np.float32)


def icosphere(f: int):
    """Returns (xyz (20 f^2, 9) float32, rgb (20 f^2, 3) uint8, scene_max float32)."""
    assert f >= 1
    n_face = f * f
    xyz = np.empty((20 * n_face, 9), np.float32)
    rgb = np.empty((20 * n_face, 3), np.uint8)
    for k, (a, b, c) in enumerate(_ICO_F):
        xyz[k * n_face:(k + 1) * n_face] = _face_soup(_ICO_V[a], _ICO_V[b], _ICO_V[c], f)
        rgb[k * n_face:(k + 1) * n_face] = FACE_COLOURS[k]
    # OBJ-style bbox fold from 0 (geometry.rs:85-88): max component of any vertex
    scene_max = np.float32(max(0.0, float(xyz.max())))
    return xyz, rgb, scene_max


def random_soup(seed: int, n: int, spread: float = 1.3, kind: str = "uniform"):
    """Fuzz inputs for the parity tests: uniform soups plus slivers, collinear and
    duplicated triangles (exact depth ties), axis-aligned edges."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-spread, spread, size=(n, 9)).astype(np.float32)
    if kind == "small":
        c = rng.uniform(-spread, spread, size=(n, 1, 3))
        xyz = (c + rng.uniform(-0.08, 0.08, size=(n, 3, 3))).reshape(n, 9).astype(np.float32)
    elif kind == "sliver":
        xyz[:, 3:6] = xyz[:, 0:3] + rng.uniform(-1e-6, 1e-6, size=(n, 3)).astype(np.float32)
    elif kind == "collinear":
        t = rng.uniform(0, 1, size=(n, 1)).astype(np.float32)
        xyz[:, 6:9] = xyz[:, 0:3] + t * (xyz[:, 3:6] - xyz[:, 0:3])
    elif kind == "dup":
        half = n // 2
        xyz[half:2 * half] = xyz[:half]
    elif kind == "axis":
        xyz = np.round(xyz * 4) / 4
    rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    scene_max = np.float32(max(0.0, float(xyz.max()))) if kind != "offscreen" else np.float32(1.0)
    if scene_max == 0:
        scene_max = np.float32(1.0)
    return xyz.astype(np.float32), rgb, scene_max
