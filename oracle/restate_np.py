"""Second, independent restatement of the reference raster path in numpy float32 --
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Written from the reference source (src/rasterizer.rs:48-93, src/geometry.rs:37-56,
src/context.rs:93-141) without looking at sloth_oracle.c's loop structure: one Python
iteration per triangle and per row, candidates of a row evaluated as float32 vectors.
numpy float32 arithmetic rounds every operation to binary32 and never fuses, which is
what rustc emits.  tests/test_oracle.py requires this file and sloth_oracle.c to agree
bit for bit; that agreement (plus the reference's three unit-test vectors and the
candidate/covered/z-write counts of SURVEY.md Appendix D, produced by a third
throwaway restatement during the survey) is the only pin available without a Rust
toolchain -- PARITY UNPINNED against a real `sloth` binary.
"""
from __future__ import annotations

import numpy as np

F = np.float32
F32_MAX = F(3.40282347e38)
THR = np.array([0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.90, 1.0], F)
GLYPH = np.frombuffer(b".:-=+*#%@ ", np.uint8)


def _sat_usize(v) -> int:
    v = float(v)
    if not v > 0.0:  # negative, zero, NaN
        return 0
    return int(min(v, 2.0 ** 64 - 1))


def matmul44(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """4x4 (row, col) arrays; nalgebra order: per element ((a0*b0 + a1*b1) + a2*b2) + a3*b3."""
    C = np.zeros((4, 4), F)
    for i in range(4):
        for j in range(4):
            acc = F(A[i, 0] * B[0, j])
            for k in range(1, 4):
                acc = F(acc + F(A[i, k] * B[k, j]))
            C[i, j] = acc
    return C


def utransform(W: int, H: int, scale0) -> np.ndarray:
    w16, h16 = W & 0xFFFF, H & 0xFFFF
    T = np.eye(4, dtype=F)
    if w16 == 0 and h16 == 0:
        return T
    fw, fh = F(w16), F(h16)
    with np.errstate(all="ignore"):
        s = F(F(np.fmin(fh, F(fw / F(2.0))) / F(scale0)) / F(2.0))
    T[0, 0], T[0, 3] = s, F(fw / F(4.0))
    T[1, 1], T[1, 3] = F(-s), F(fh / F(2.0))
    T[2, 2] = s
    return T


def render(xyz, rgb, scale0, W, H, rot_colmajor, image=True):
    xyz = np.asarray(xyz, F).reshape(-1, 3, 3)
    rgb = np.asarray(rgb, np.uint8).reshape(-1, 3)
    R = np.asarray(rot_colmajor, F).reshape(4, 4).T  # column-major flat -> (row, col)
    with np.errstate(all="ignore"):
        M = matmul44(utransform(W, H, scale0), R)
        cells = np.full(W * H + (H if image else 0), ord(" "), np.uint32)
        zbuf = np.full(W * H, F32_MAX, F)
        for t in range(xyz.shape[0]):
            v = np.empty((3, 4), F)
            for k in range(3):
                p = np.array([xyz[t, k, 0], xyz[t, k, 1], xyz[t, k, 2], 1.0], F)
                for i in range(4):
                    acc = F(M[i, 0] * p[0])
                    for j in range(1, 4):
                        acc = F(acc + F(M[i, j] * p[j]))
                    v[k, i] = acc
            v1, v2, v3 = v
            mn = np.fmin(v1, np.fmin(v2, v3))
            mx = np.fmax(v1, np.fmax(v2, v3))
            minx = _sat_usize(np.ceil(np.fmax(mn[0], F(1.0))))
            miny = _sat_usize(np.ceil(np.fmax(mn[1], F(1.0))))
            maxx = _sat_usize(np.ceil(np.fmin(F(mx[0] * F(2.0)), F(W - 1))))
            maxy = _sat_usize(np.ceil(np.fmin(mx[1], F(H - 1))))

            def orient(a, b, cx, cy):
                return F(b[0] - a[0]) * (cy - a[1]) - F(b[1] - a[1]) * (cx - a[0])

            area = F(F(F(v2[0] - v1[0]) * F(v3[1] - v1[1])) - F(F(v2[1] - v1[1]) * F(v3[0] - v1[0])))
            a = F(F(1.0) / area)
            # Triangle::normal().z
            e1, e2 = (v2 - v1).astype(F), (v3 - v1).astype(F)
            nx = F(F(e1[1] * e2[2]) - F(e1[2] * e2[1]))
            ny = F(F(e1[2] * e2[0]) - F(e1[0] * e2[2]))
            nz = F(F(e1[0] * e2[1]) - F(e1[1] * e2[0]))
            n2 = F(F(F(nx * nx) + F(nz * nz)) + F(F(ny * ny) + F(F(0.0) * F(0.0))))
            nzu = F(nz / F(np.sqrt(n2)))
            cell = np.uint32(0) | (np.uint32(rgb[t, 0]) << 8) | (np.uint32(rgb[t, 1]) << 16) | (np.uint32(rgb[t, 2]) << 24)
            for y in range(miny, maxy):
                if maxx > minx:
                    px = np.arange(minx, maxx, dtype=np.int64).astype(F)
                    py = F(y)
                    w0 = orient(v2, v3, px, py).astype(F)
                    w1 = orient(v3, v1, px, py).astype(F)
                    w2 = orient(v1, v2, px, py).astype(F)
                    cov = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
                    if cov.any():
                        shade = (F(nzu * a) * ((w0 + w1).astype(F) + w2).astype(F)).astype(F)
                        z = (v1[2] + (a * ((w1 * F(v2[2] - v1[2])).astype(F) + (w2 * F(v3[2] - v1[2])).astype(F)).astype(F)).astype(F)).astype(F)
                        ids = y * W + 2 * np.arange(minx, maxx, dtype=np.int64)
                        win = cov & (z < zbuf[ids])   # ids of one row are distinct
                        if win.any():
                            g = np.full(px.shape, 9, np.int64)
                            for i in range(8, -1, -1):
                                g = np.where(shade <= THR[i], i, g)
                            wi = ids[win]
                            zbuf[wi] = z[win]
                            pix = cell | GLYPH[g[win]].astype(np.uint32)
                            cells[wi] = pix
                            cells[wi + 1] = pix
                if image:
                    cells[y * W + 1] = ord("\n")
    return cells, zbuf
