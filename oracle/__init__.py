"""CPU oracle for the rust-sloth raster path -- TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/sloth_oracle.c`` (the plain-C restatement of
``src/rasterizer.rs``, ``src/geometry.rs``, ``src/context.rs`` of the
reference).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package; the product library never does.

PARITY UNPINNED against a real ``sloth`` binary (no Rust toolchain here); see
the header of ``sloth_oracle.c`` and DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsloth_oracle.so")
_lib = None

F32_MAX = np.float32(3.40282347e38)
BLANK_CELL = np.uint32(ord(" "))
NEWLINE_CELL = np.uint32(ord("\n"))


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no FMA contraction); returns the .so path."""
    src = os.path.join(_HERE, "sloth_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libsloth_oracle.so"])
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.oracle_mat4_mul.argtypes = [fp, fp, fp]
        L.oracle_utransform.argtypes = [C.c_uint32, C.c_uint32, C.c_float, fp]
        L.oracle_utransform.restype = C.c_int
        L.oracle_rotation.argtypes = [C.c_float, C.c_float, C.c_float, fp]
        L.oracle_turntable.argtypes = [C.c_float, C.c_uint32, fp, C.c_size_t]
        L.oracle_turntable.restype = C.c_size_t
        L.oracle_render.argtypes = [
            fp, C.POINTER(C.c_uint8), C.c_size_t, C.c_float, C.c_uint32, C.c_uint32, C.c_int, fp,
            C.c_int, C.c_size_t, C.c_size_t, fp, C.c_char_p, C.POINTER(C.c_uint32), fp,
            C.POINTER(C.c_uint64)]
        L.oracle_render.restype = C.c_int
        for name in ("oracle_triangle_aabb", "oracle_triangle_normal", "oracle_vec4_normalize"):
            getattr(L, name).restype = None
        L.oracle_triangle_aabb.argtypes = [fp, fp, fp]
        L.oracle_triangle_mul.argtypes = [fp, fp]
        L.oracle_triangle_normal.argtypes = [fp, fp]
        L.oracle_vec4_normalize.argtypes = [fp, fp]
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def mat4_mul(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Column-major flat 16-vectors in, column-major flat out."""
    A = np.ascontiguousarray(A, np.float32).reshape(16)
    B = np.ascontiguousarray(B, np.float32).reshape(16)
    out = np.empty(16, np.float32)
    lib().oracle_mat4_mul(_fp(A), _fp(B), _fp(out))
    return out


def utransform(W: int, H: int, scale0: float) -> np.ndarray:
    out = np.zeros(16, np.float32)
    out[[0, 5, 10, 15]] = 1.0
    lib().oracle_utransform(W, H, np.float32(scale0), _fp(out))
    return out


def rotation(roll: float, pitch: float, yaw: float) -> np.ndarray:
    out = np.empty(16, np.float32)
    lib().oracle_rotation(np.float32(roll), np.float32(pitch), np.float32(yaw), _fp(out))
    return out


def turntable(y_arg: float, n_frames: int) -> np.ndarray:
    cap = max(int(n_frames), 1) + 4
    buf = np.empty(cap, np.float32)
    n = lib().oracle_turntable(np.float32(y_arg), int(n_frames), _fp(buf), cap)
    return buf[:n].copy()


def render(xyz: np.ndarray, rgb: np.ndarray, scale0: float, W: int, H: int, rot: np.ndarray,
           image: bool = True, mode: int = 0, tri_first: int = 0, tri_step: int = 1,
           thr: np.ndarray | None = None, glyph: bytes | None = None):
    """Render one frame.  Returns (cells uint32[W*H(+H)], zbuf f32[W*H], counters dict)."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 9)
    n = xyz.shape[0]
    rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
    assert rgb.shape[0] == n
    rot = np.ascontiguousarray(rot, np.float32).reshape(16)
    cells = np.empty(W * H + (H if image else 0), np.uint32)
    zbuf = np.empty(W * H, np.float32)
    cnt = np.zeros(4, np.uint64)
    thr_p = None
    if thr is not None:
        thr = np.ascontiguousarray(thr, np.float32).reshape(9)
        thr_p = _fp(thr)
    if glyph is not None:
        assert len(glyph) == 10
    rc = lib().oracle_render(
        _fp(xyz), rgb.ctypes.data_as(C.POINTER(C.c_uint8)), n, np.float32(scale0), W, H,
        1 if image else 0, _fp(rot), mode, tri_first, tri_step, thr_p, glyph,
        cells.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(zbuf),
        cnt.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0
    counters = {"candidates": int(cnt[0]), "covered": int(cnt[1]), "zwrites": int(cnt[2]),
                "stamps": int(cnt[3])}
    return cells, zbuf, counters


def cells_to_text(cells: np.ndarray) -> str:
    """Context::flush in no-colour mode (src/context.rs:59-62): glyphs concatenated."""
    return bytes((cells & 0xFF).astype(np.uint8)).decode("latin-1")
