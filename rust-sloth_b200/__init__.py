"""rust_sloth_b200 -- B200-native raster path of ecumene/rust-sloth.

Host-side mirror (Python over ctypes) of the reference's public items on the
raster path, bound to the C ABI in ``include/sloth_b200.h``:

==========================  ===========================================
reference (Rust)            here
==========================  ===========================================
``Context::blank``          :meth:`Context.blank`            context.rs:22
``match_dimensions``        ``ctx.width / ctx.height``       inputs.rs:159
``Context::update``         :meth:`Context.update`           context.rs:93
``Context::clear``          :meth:`Context.clear`            context.rs:35
``draw_mesh``               :func:`draw_mesh`                rasterizer.rs:39
``Context.frame_buffer``    :attr:`Context.frame_buffer`     context.rs:16
``Context.z_buffer``        :attr:`Context.z_buffer`         context.rs:17
``Context::flush``          :meth:`Context.flush`            context.rs:50
``default_shader``          :func:`default_shader`           rasterizer.rs:5
``match_meshes``            :func:`match_meshes`             inputs.rs:95
``match_turntable`` + loop  :func:`turntable_pitches`        inputs.rs:131, main.rs:55-106
==========================  ===========================================

The directory is called ``rust-sloth_b200`` (not importable as written); the
root-level ``rust_sloth_b200.py`` shim loads it under this legal name.

There is no CPU fallback: importing works anywhere, but creating a
:class:`Context` raises :class:`SlothError` when ``libsloth_b200.so`` is
missing or no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLOTH_B200_LIB") or os.path.join(_HERE, "libsloth_b200.so")   # override: A/B builds while profiling
HOST_LIB_PATH = os.path.join(_HERE, "libsloth_host.so")

SLOTH_OK, SLOTH_E_ARG, SLOTH_E_CUDA, SLOTH_E_STATE, SLOTH_E_TOO_LARGE = 0, -1, -2, -3, -4
SLOTH_E_IO, SLOTH_E_PARSE, SLOTH_E_UNSUPPORTED = -5, -6, -7
PATH_AUTO, PATH_SOUP, PATH_INDEXED = 0, 1, 2   # sloth_ctx_set_path
WIRE_CELLS, WIRE_SPANS = 0, 1                  # sloth_ctx_set_wire

# every symbol include/sloth_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "sloth_ctx_create", "sloth_ctx_destroy", "sloth_scene_set", "sloth_scene_set_indexed", "sloth_ctx_set_path",
    "sloth_ctx_resize", "sloth_render",
    "sloth_render_batch", "sloth_render_device", "sloth_ctx_sync", "sloth_ctx_set_band",
    "sloth_shader_set", "sloth_stats_get", "sloth_stats_enable", "sloth_last_error",
    "sloth_rotation_from_euler", "sloth_utransform", "sloth_turntable_pitches", "sloth_cells_per_frame",
    "sloth_pinned_alloc", "sloth_pinned_free", "sloth_host_register", "sloth_host_unregister", "sloth_ctx_stream",
    "sloth_render_device_batch",
    "sloth_text_capacity", "sloth_render_text", "sloth_render_text_batch", "sloth_flush_device",
    "sloth_scene_load", "sloth_loader_begin", "sloth_loader_add_obj", "sloth_loader_add_stl", "sloth_loader_commit",
    "sloth_scene_size", "sloth_scene_get",
    "sloth_device_alloc", "sloth_device_free", "sloth_ipc_export", "sloth_ipc_open", "sloth_ipc_close", "sloth_device_read", "sloth_device_write",
    "sloth_ctx_set_wire", "sloth_wire_stats", "sloth_expand_spans",
]


class SlothError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sloth_b200 error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [
        ("frames", C.c_uint64), ("kernel_launches", C.c_uint64), ("fragments", C.c_uint64),
        ("n_tri", C.c_uint32), ("walk_tris", C.c_uint32), ("walk_items", C.c_uint32),
        ("irregular_tris", C.c_uint32), ("stamp_fixups", C.c_uint32),
        ("last_frame_ms", C.c_float), ("geom_ms", C.c_float), ("walk_ms", C.c_float),
        ("resolve_ms", C.c_float),
        ("chunks_processed", C.c_uint32),
        ("load_read_ms", C.c_float), ("load_parse_ms", C.c_float), ("load_commit_ms", C.c_float),
        ("n_vert", C.c_uint32), ("geom_path", C.c_uint32), ("xform_ms", C.c_float),
        ("l2_window_bytes", C.c_uint64), ("l2_persist_max", C.c_uint64),
        ("tile_tris", C.c_uint32), ("tile_pairs", C.c_uint32), ("tiles_used", C.c_uint32),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None
_host = None


def load_library() -> C.CDLL:
    """dlopen the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SlothError(SLOTH_E_CUDA, f"{LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    fp, vp = C.POINTER(C.c_float), C.c_void_p
    L.sloth_ctx_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.sloth_ctx_destroy.argtypes = [vp]
    L.sloth_scene_set.argtypes = [vp, fp, C.POINTER(C.c_uint8), C.c_size_t, C.c_float]
    L.sloth_scene_set_indexed.argtypes = [vp, fp, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.c_size_t, C.c_float]
    L.sloth_ctx_set_path.argtypes = [vp, C.c_int]
    L.sloth_ctx_resize.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.sloth_render.argtypes = [vp, fp, C.POINTER(C.c_uint32), fp]
    L.sloth_render_batch.argtypes = [vp, fp, C.c_size_t, C.POINTER(C.c_uint32)]
    L.sloth_render_device.argtypes = [vp, fp, vp]
    L.sloth_render_device_batch.argtypes = [vp, fp, C.c_size_t, vp, C.c_size_t]
    L.sloth_text_capacity.argtypes = [vp, C.c_int]
    L.sloth_text_capacity.restype = C.c_size_t
    L.sloth_render_text.argtypes = [vp, fp, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.sloth_render_text_batch.argtypes = [vp, fp, C.c_size_t, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.sloth_flush_device.argtypes = [vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.sloth_ctx_sync.argtypes = [vp]
    L.sloth_ctx_stream.argtypes = [vp]
    L.sloth_ctx_stream.restype = vp
    L.sloth_ctx_set_band.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.sloth_shader_set.argtypes = [vp, fp, C.c_char_p]
    L.sloth_stats_get.argtypes = [vp, C.POINTER(Stats)]
    L.sloth_stats_enable.argtypes = [vp, C.c_uint32]
    L.sloth_last_error.restype = C.c_char_p
    L.sloth_rotation_from_euler.argtypes = [C.c_float, C.c_float, C.c_float, fp]
    L.sloth_rotation_from_euler.restype = None
    L.sloth_utransform.argtypes = [C.c_uint32, C.c_uint32, C.c_float, fp]
    L.sloth_utransform.restype = None
    L.sloth_turntable_pitches.argtypes = [C.c_float, C.c_uint32, fp, C.c_size_t]
    L.sloth_turntable_pitches.restype = C.c_size_t
    L.sloth_cells_per_frame.argtypes = [vp]
    L.sloth_cells_per_frame.restype = C.c_size_t
    L.sloth_pinned_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.sloth_pinned_free.argtypes = [vp]
    L.sloth_host_register.argtypes = [vp, C.c_size_t]
    L.sloth_host_unregister.argtypes = [vp]
    L.sloth_scene_load.argtypes = [vp, C.c_char_p, C.POINTER(C.c_size_t), fp]
    L.sloth_loader_begin.argtypes = [vp]
    L.sloth_loader_add_obj.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_char_p]
    L.sloth_loader_add_stl.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.sloth_loader_commit.argtypes = [vp, C.POINTER(C.c_size_t), fp]
    L.sloth_scene_size.argtypes = [vp]
    L.sloth_scene_size.restype = C.c_size_t
    L.sloth_scene_get.argtypes = [vp, fp, C.POINTER(C.c_uint8), fp]
    L.sloth_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp)]
    L.sloth_device_free.argtypes = [C.c_int, vp]
    L.sloth_ipc_export.argtypes = [C.c_int, vp, C.c_char_p]
    L.sloth_ipc_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.sloth_ipc_close.argtypes = [C.c_int, vp]
    L.sloth_device_read.argtypes = [vp, vp, vp, C.c_size_t]
    L.sloth_device_write.argtypes = [vp, vp, vp, C.c_size_t]
    L.sloth_ctx_set_wire.argtypes = [vp, C.c_int]
    L.sloth_wire_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.sloth_expand_spans.argtypes = [C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t]
    _lib = L
    return L


IPC_HANDLE_BYTES = 64


def host_register(array: np.ndarray) -> None:
    """Page-lock memory the caller owns (e.g. a shared-memory frame mapped by every per-GPU process)."""
    _check(load_library().sloth_host_register(C.c_void_p(array.ctypes.data), array.nbytes))


def host_unregister(array: np.ndarray) -> None:
    _check(load_library().sloth_host_unregister(C.c_void_p(array.ctypes.data)))


def device_alloc(device: int, nbytes: int) -> int:
    p = C.c_void_p()
    _check(load_library().sloth_device_alloc(int(device), int(nbytes), C.byref(p)))
    return int(p.value)


def device_free(device: int, ptr: int) -> None:
    _check(load_library().sloth_device_free(int(device), C.c_void_p(ptr)))


def ipc_export(device: int, ptr: int) -> bytes:
    buf = C.create_string_buffer(IPC_HANDLE_BYTES)
    _check(load_library().sloth_ipc_export(int(device), C.c_void_p(ptr), buf))
    return buf.raw


def ipc_open(device: int, handle: bytes) -> int:
    p = C.c_void_p()
    _check(load_library().sloth_ipc_open(int(device), handle, C.byref(p)))
    return int(p.value)


def ipc_close(device: int, ptr: int) -> None:
    _check(load_library().sloth_ipc_close(int(device), C.c_void_p(ptr)))


def _check(rc: int) -> None:
    if rc != SLOTH_OK:
        raise SlothError(rc, load_library().sloth_last_error().decode("utf-8", "replace"))


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


# --------------------------------------------------------------------------------------
# geometry.rs mirrors
# --------------------------------------------------------------------------------------
@dataclass
class AABB:  # geometry.rs:5-15
    min: np.ndarray
    max: np.ndarray


class SimpleMesh:
    """geometry.rs:78-81 as a soup: ``xyz`` (n,9) f32, ``rgb`` (n,3) u8, ``bounding_box``."""

    def __init__(self, xyz, rgb, bbox_min=None, bbox_max=None):
        self.xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 9)
        self.rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
        if self.xyz.shape[0] != self.rgb.shape[0]:
            raise ValueError("xyz and rgb disagree on the triangle count")
        if bbox_max is None:  # OBJ rule: the fold starts at 0 (geometry.rs:85-88)
            pts = self.xyz.reshape(-1, 3)
            zero = np.zeros(3, np.float32)
            bbox_min = np.minimum(pts.min(axis=0), zero) if len(pts) else zero
            bbox_max = np.maximum(pts.max(axis=0), zero) if len(pts) else zero
        self.bounding_box = AABB(np.asarray(bbox_min, np.float32), np.asarray(bbox_max, np.float32))

    def __len__(self) -> int:
        return self.xyz.shape[0]


def scene_scale0(meshes) -> np.float32:
    """context.rs:106-113: fold(max) of bounding_box.max.{x,y,z} over meshes, from 0.0."""
    s = np.float32(0.0)
    for m in meshes:
        for d in range(3):
            s = np.fmax(s, np.float32(m.bounding_box.max[d]))
    return np.float32(s)


class _HostScene(C.Structure):
    _fields_ = [("xyz", C.POINTER(C.c_float)), ("rgb", C.POINTER(C.c_uint8)), ("n_tri", C.c_size_t),
                ("n_meshes", C.c_size_t), ("mesh_sizes", C.POINTER(C.c_size_t)),
                ("mesh_bbox", C.POINTER(C.c_float)), ("scale0", C.c_float), ("error", C.c_char * 512)]


def _host_lib() -> C.CDLL:
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise SlothError(SLOTH_E_STATE, f"{HOST_LIB_PATH} is missing -- run __graft_entry__.build()")
        H = C.CDLL(HOST_LIB_PATH)
        H.sloth_host_load.argtypes = [C.c_char_p]
        H.sloth_host_load.restype = C.POINTER(_HostScene)
        H.sloth_host_free.argtypes = [C.POINTER(_HostScene)]
        H.sloth_host_parse_f32.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.c_size_t]
        H.sloth_host_parse_f32.restype = C.c_size_t
        _host = H
    return _host


def parse_f32_tokens(tokens) -> tuple[np.ndarray, np.ndarray]:
    """The device loader's decimal -> f32 routine (csrc/dec_float.cuh) compiled for the host: (values, status)
    with status 0 = ok, 1 = not a number, 2 = not decided (see loader.cuh).  For the CPU tests."""
    text = "\n".join(tokens).encode()
    out = np.zeros(len(tokens), np.float32)
    status = np.zeros(len(tokens), np.uint8)
    n = _host_lib().sloth_host_parse_f32(text, len(text), out.ctypes.data_as(C.POINTER(C.c_float)),
                                         status.ctypes.data_as(C.POINTER(C.c_uint8)), len(tokens))
    assert n == len(tokens), "empty tokens are not allowed"
    return out, status


def match_meshes(arg: str) -> list[SimpleMesh]:
    """inputs.rs:95-129: one CLI value, split on ' ', OBJ via tobj rules, STL via stl_io rules."""
    H = _host_lib()
    sp = H.sloth_host_load(arg.encode())
    try:
        s = sp.contents
        if s.error:
            raise SlothError(SLOTH_E_ARG, s.error.decode("utf-8", "replace"))
        n = s.n_tri
        xyz = np.ctypeslib.as_array(s.xyz, shape=(n * 9,)).copy().reshape(n, 9) if n else np.zeros((0, 9), np.float32)
        rgb = np.ctypeslib.as_array(s.rgb, shape=(n * 3,)).copy().reshape(n, 3) if n else np.zeros((0, 3), np.uint8)
        meshes, off = [], 0
        for i in range(s.n_meshes):
            k = s.mesh_sizes[i]
            bb = [s.mesh_bbox[i * 6 + d] for d in range(6)]
            meshes.append(SimpleMesh(xyz[off:off + k], rgb[off:off + k], bb[:3], bb[3:]))
            off += k
        return meshes
    finally:
        H.sloth_host_free(sp)


# --------------------------------------------------------------------------------------
# rasterizer.rs / main.rs mirrors
# --------------------------------------------------------------------------------------
DEFAULT_THRESHOLDS = np.array([0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.90, 1.0], np.float32)
DEFAULT_GLYPHS = b".:-=+*#%@ "


def default_shader(shade) -> str:
    """rasterizer.rs:5-27 (host copy; the device evaluates the same table per fragment)."""
    shade = np.float32(shade)
    for t, g in zip(DEFAULT_THRESHOLDS, DEFAULT_GLYPHS[:9]):
        if shade <= t:
            return chr(g)
    return " "


def rotation_from_euler(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """Rotation3::from_euler_angles(..).to_homogeneous(), main.rs:76-77; column-major float32[16]."""
    out = np.empty(16, np.float32)
    load_library().sloth_rotation_from_euler(np.float32(roll), np.float32(pitch), np.float32(yaw), _fp(out))
    return out


def utransform(width: int, height: int, scene_max: float) -> np.ndarray:
    out = np.empty(16, np.float32)
    load_library().sloth_utransform(width, height, np.float32(scene_max), _fp(out))
    return out


def turntable_pitches(y_arg: float, n_frames: int) -> np.ndarray:
    """Pitch of every frame `image -j N` renders (inputs.rs:131-149, main.rs:55-58,92-106)."""
    L = load_library()
    cap = max(int(n_frames), 1) + 4
    buf = np.empty(cap, np.float32)
    n = L.sloth_turntable_pitches(np.float32(y_arg), int(n_frames), _fp(buf), cap)
    return buf[:n].copy()


class PinnedBuffer:
    """Page-locked host memory for the frame buffers of a batch (cudaHostAlloc)."""

    def __init__(self, n_cells: int):
        self._ptr = C.c_void_p()
        self.n = int(n_cells)
        _check(load_library().sloth_pinned_alloc(max(self.n, 1) * 4, C.byref(self._ptr)))
        self.array = np.ctypeslib.as_array(C.cast(self._ptr, C.POINTER(C.c_uint32)), shape=(max(self.n, 1),))[:self.n]

    def free(self):
        if self._ptr:
            load_library().sloth_pinned_free(self._ptr)
            self._ptr = C.c_void_p()
            self.array = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """context.rs:12-19.  Owns one GPU context; not thread-safe (one per thread / GPU)."""

    def __init__(self, image: bool, device: int = 0, path: int | None = None):
        self._L = load_library()
        self._h = C.c_void_p()
        _check(self._L.sloth_ctx_create(device, 1 if image else 0, C.byref(self._h)))
        if path is not None:
            _check(self._L.sloth_ctx_set_path(self._h, int(path)))
        self.image = bool(image)
        self.device = device
        self.width = 0
        self.height = 0
        self.utransform = np.eye(4, dtype=np.float32).T.reshape(16).copy()
        self._sized = (0, 0)
        self._scene_id = None
        self._scene_max = np.float32(0)
        self._band = (0, 0)
        self._pending_rot = None
        self._frame = None
        self._z = None

    @classmethod
    def blank(cls, image: bool, device: int = 0, path: int | None = None) -> "Context":
        return cls(image, device, path)

    def set_path(self, path: int) -> None:
        """PATH_AUTO / PATH_SOUP / PATH_INDEXED for the scenes set from now on (sloth_ctx_set_path)."""
        _check(self._L.sloth_ctx_set_path(self._h, int(path)))

    def set_wire(self, wire: int) -> None:
        """WIRE_CELLS (plain 4-byte cells over PCIe) or WIRE_SPANS (run lists, cells rebuilt by host threads inside the
        library) for render / render_batch from now on (sloth_ctx_set_wire); resets wire_stats()."""
        _check(self._L.sloth_ctx_set_wire(self._h, int(wire)))

    def wire_stats(self) -> dict:
        out = (C.c_uint64 * 4)()
        _check(self._L.sloth_wire_stats(self._h, out))
        return {"frames": int(out[0]), "plain_frames": int(out[1]), "d2h_bytes": int(out[2]), "threads": int(out[3])}

    def close(self):
        if self._h:
            self._L.sloth_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene / size ------------------------------------------------------------------
    def set_scene(self, xyz: np.ndarray, rgb: np.ndarray, scene_max: float) -> None:
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 9)
        rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
        _check(self._L.sloth_scene_set(self._h, _fp(xyz), rgb.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       xyz.shape[0], np.float32(scene_max)))
        self._scene_max = np.float32(scene_max)
        self.n_tri = xyz.shape[0]

    def set_scene_indexed(self, positions: np.ndarray, indices: np.ndarray, rgb: np.ndarray, scene_max: float) -> None:
        """The mesh queue before de-indexing (geometry.rs:99-107): positions (v,3) f32, indices (n,3) u32, rgb (n,3) u8."""
        positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        indices = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
        assert rgb.shape[0] == indices.shape[0]
        _check(self._L.sloth_scene_set_indexed(self._h, _fp(positions), positions.shape[0],
                                               indices.ctypes.data_as(C.POINTER(C.c_uint32)),
                                               rgb.ctypes.data_as(C.POINTER(C.c_uint8)), indices.shape[0], np.float32(scene_max)))
        self._scene_max = np.float32(scene_max)
        self.n_tri = indices.shape[0]

    # -- device loaders (inputs.rs:95-129 + geometry.rs:83-189 on the GPU) ----------------
    def _loaded(self, n: C.c_size_t, m: C.c_float) -> tuple[int, np.float32]:
        self.n_tri = int(n.value)
        self._scene_max = np.float32(m.value)
        return self.n_tri, self._scene_max

    def load_models(self, arg: str) -> tuple[int, np.float32]:
        """match_meshes for one CLI value, parsed on the device; returns (triangles, scene_max)."""
        n, m = C.c_size_t(0), C.c_float(0.0)
        _check(self._L.sloth_scene_load(self._h, arg.encode(), C.byref(n), C.byref(m)))
        return self._loaded(n, m)

    def load_bytes(self, files) -> tuple[int, np.float32]:
        """files: iterable of ("obj", text_bytes, mtl_dir) / ("stl", bytes); draw order = list order."""
        _check(self._L.sloth_loader_begin(self._h))
        for f in files:
            if f[0] == "obj":
                mtl_dir = f[2] if len(f) > 2 and f[2] is not None else ""
                _check(self._L.sloth_loader_add_obj(self._h, f[1], len(f[1]), mtl_dir.encode()))
            elif f[0] == "stl":
                _check(self._L.sloth_loader_add_stl(self._h, f[1], len(f[1])))
            else:
                raise ValueError(f"unknown model kind {f[0]!r}")
        n, m = C.c_size_t(0), C.c_float(0.0)
        _check(self._L.sloth_loader_commit(self._h, C.byref(n), C.byref(m)))
        return self._loaded(n, m)

    def read_device(self, ptr: int, n_cells: int) -> np.ndarray:
        """n_cells uint32 cells from device memory (own or IPC-mapped) to the host."""
        out = np.empty(int(n_cells), np.uint32)
        _check(self._L.sloth_device_read(self._h, C.c_void_p(ptr), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def write_device(self, ptr: int, cells: np.ndarray) -> None:
        cells = np.ascontiguousarray(cells, np.uint32)
        _check(self._L.sloth_device_write(self._h, C.c_void_p(ptr), cells.ctypes.data_as(C.c_void_p), cells.nbytes))

    def scene(self) -> tuple[np.ndarray, np.ndarray, np.float32]:
        """The resident soup read back from the device: (xyz [n,9] f32, rgb [n,3] u8, scene_max)."""
        n = int(self._L.sloth_scene_size(self._h))
        xyz = np.empty((n, 9), np.float32)
        rgb = np.empty((n, 3), np.uint8)
        m = C.c_float(0.0)
        _check(self._L.sloth_scene_get(self._h, _fp(xyz), rgb.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(m)))
        return xyz, rgb, np.float32(m.value)

    def resize(self, width: int, height: int) -> None:
        _check(self._L.sloth_ctx_resize(self._h, int(width), int(height)))
        self.width, self.height = int(width), int(height)
        self._sized = (self.width, self.height)
        self._band = (0, 0)

    def set_band(self, row0: int, row1: int) -> None:
        _check(self._L.sloth_ctx_set_band(self._h, int(row0), int(row1)))
        self._band = (int(row0), int(row1))

    def update(self, old_size, meshes) -> None:
        """Context::update (context.rs:93-141): adopts the size and the scene scale.

        The mesh queue is uploaded the first time (or when a different list is passed)."""
        del old_size  # the reference passes a by-value (0,0) every frame (main.rs:53,78)
        if (self.width, self.height) != self._sized:
            self.resize(self.width, self.height)
        key = tuple(id(m) for m in meshes)
        if key != self._scene_id:
            xyz = np.concatenate([m.xyz for m in meshes]) if meshes else np.zeros((0, 9), np.float32)
            rgb = np.concatenate([m.rgb for m in meshes]) if meshes else np.zeros((0, 3), np.uint8)
            self.set_scene(xyz, rgb, scene_scale0(meshes))
            self._scene_id = key
            self._meshes = list(meshes)
        self.utransform = utransform(self.width, self.height, self._scene_max)

    def clear(self) -> None:
        """Context::clear (context.rs:35-45): starts a new frame."""
        self._pending_rot = None
        self._drawn = []
        self._frame = None
        self._z = None

    def _draw(self, mesh, transform) -> None:
        rot = np.ascontiguousarray(transform, np.float32).reshape(16)
        if self._pending_rot is not None and not np.array_equal(self._pending_rot.view(np.uint32), rot.view(np.uint32)):
            raise NotImplementedError("all draw_mesh calls of one frame must use the same transform "
                                      "(the reference's only caller does, main.rs:80-83)")
        self._pending_rot = rot
        self._drawn.append(mesh)
        self._frame = None

    def _resolve(self, want_z: bool = False) -> None:
        if self._frame is not None and (self._z is not None or not want_z):
            return
        drawn = getattr(self, "_drawn", [])
        if [id(m) for m in drawn] != [id(m) for m in getattr(self, "_meshes", [])]:
            raise NotImplementedError("a frame must draw exactly the mesh queue passed to update(), in order")
        if self._pending_rot is None:
            raise SlothError(SLOTH_E_STATE, "nothing drawn since clear()")
        self._frame, self._z = self.render(self._pending_rot, want_z=want_z)

    @property
    def frame_buffer(self) -> np.ndarray:
        """uint32 cells: glyph | r<<8 | g<<16 | b<<24 (the reference's Vec<(char,(u8,u8,u8))>)."""
        self._resolve()
        return self._frame

    @property
    def z_buffer(self) -> np.ndarray:
        self._resolve(want_z=True)
        return self._z

    # -- direct entry points -----------------------------------------------------------
    def cells_per_frame(self) -> int:
        return int(self._L.sloth_cells_per_frame(self._h))

    def render_into(self, rot: np.ndarray, out: np.ndarray) -> None:
        """sloth_render straight into caller memory (cells_per_frame() uint32 cells, ideally page-locked)."""
        rot = np.ascontiguousarray(rot, np.float32).reshape(16)
        assert out.dtype == np.uint32 and out.flags.c_contiguous and out.size >= self.cells_per_frame()
        _check(self._L.sloth_render(self._h, _fp(rot), out.ctypes.data_as(C.POINTER(C.c_uint32)), None))

    def render(self, rot: np.ndarray, want_z: bool = False):
        rot = np.ascontiguousarray(rot, np.float32).reshape(16)
        cells = np.empty(self.cells_per_frame(), np.uint32)
        z = np.empty(self.width * self.height, np.float32) if want_z else None
        _check(self._L.sloth_render(self._h, _fp(rot), cells.ctypes.data_as(C.POINTER(C.c_uint32)),
                                    _fp(z) if want_z else None))
        return cells, z

    def render_batch(self, rots: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        rots = np.ascontiguousarray(rots, np.float32).reshape(-1, 16)
        n = rots.shape[0]
        cpf = self.cells_per_frame()
        if out is None:
            out = np.empty(n * cpf, np.uint32)
        assert out.dtype == np.uint32 and out.size >= n * cpf and out.flags.c_contiguous
        _check(self._L.sloth_render_batch(self._h, _fp(rots), n, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out[:n * cpf].reshape(n, cpf)

    def render_device(self, rot: np.ndarray, device_ptr: int) -> None:
        rot = np.ascontiguousarray(rot, np.float32).reshape(16)
        _check(self._L.sloth_render_device(self._h, _fp(rot), C.c_void_p(device_ptr)))

    def render_device_batch(self, rots: np.ndarray, device_ptr: int, frame_stride_cells: int = 0) -> None:
        """Frames stay on the device (frame k at device_ptr + 4*k*frame_stride_cells); geometry of frame
        k+1 overlaps the resolve of frame k."""
        rots = np.ascontiguousarray(rots, np.float32).reshape(-1, 16)
        _check(self._L.sloth_render_device_batch(self._h, _fp(rots), rots.shape[0], C.c_void_p(device_ptr),
                                                 int(frame_stride_cells)))

    def render_text_batch(self, rots: np.ndarray, mode: int) -> list[bytes]:
        """Render + device-side Context::flush: the exact cell byte stream of every frame
        (mode 0 plain, 1 ANSI, 2 webify <span>s)."""
        rots = np.ascontiguousarray(rots, np.float32).reshape(-1, 16)
        n = rots.shape[0]
        cap = int(self._L.sloth_text_capacity(self._h, mode))
        stride = (cap + 63) & ~63
        buf = PinnedBuffer((n * stride + 3) // 4)
        lens = (C.c_size_t * n)()
        try:
            _check(self._L.sloth_render_text_batch(self._h, _fp(rots), n, mode, C.c_void_p(buf.array.ctypes.data), stride, lens))
            raw = buf.array.view(np.uint8)
            return [bytes(raw[k * stride:k * stride + lens[k]]) for k in range(n)]
        finally:
            buf.free()

    def stream_ptr(self) -> int:
        """cudaStream_t of this context (for torch.cuda.ExternalStream / event timing)."""
        return int(self._L.sloth_ctx_stream(self._h) or 0)

    def sync(self) -> None:
        _check(self._L.sloth_ctx_sync(self._h))

    def set_shader(self, thresholds=None, glyphs: bytes | None = None) -> None:
        if thresholds is None or glyphs is None:
            _check(self._L.sloth_shader_set(self._h, None, None))
            return
        thr = np.ascontiguousarray(thresholds, np.float32).reshape(9)
        assert len(glyphs) == 10
        _check(self._L.sloth_shader_set(self._h, _fp(thr), glyphs))

    def stats_enable(self, count_fragments: bool = False, kernel_timing: bool = False) -> None:
        _check(self._L.sloth_stats_enable(self._h, (1 if count_fragments else 0) | (2 if kernel_timing else 0)))

    def stats(self) -> dict:
        st = Stats()
        _check(self._L.sloth_stats_get(self._h, C.byref(st)))
        return st.as_dict()

    # -- presentation (host) -----------------------------------------------------------
    def flush(self, color: bool, webify: bool) -> bytes:
        """Context::flush (context.rs:50-92) as bytes instead of stdout writes."""
        return flush_bytes(self.frame_buffer, color, webify, self.image)


def draw_mesh(context: Context, mesh: SimpleMesh, transform, shader=default_shader) -> None:
    """rasterizer.rs:39-46.  Only ``default_shader`` (or a table set with
    ``Context.set_shader``) can run on the device."""
    if shader is not default_shader:
        raise NotImplementedError("arbitrary closures cannot run on the device; use Context.set_shader(thresholds, glyphs)")
    context._draw(mesh, transform)


def expand_spans(runs: np.ndarray, n_cells: int) -> np.ndarray:
    """sloth_expand_spans: the host half of the span wire format on its own.  runs: (n, 2) uint32 rows (start, cell)."""
    runs = np.ascontiguousarray(runs, np.uint32).reshape(-1, 2)
    out = np.empty(int(n_cells), np.uint32)
    _check(load_library().sloth_expand_spans(runs.ctypes.data_as(C.POINTER(C.c_uint32)), runs.shape[0],
                                    out.ctypes.data_as(C.POINTER(C.c_uint32)), int(n_cells)))
    return out


def flush_bytes(cells: np.ndarray, color: bool, webify: bool, image: bool) -> bytes:
    """Byte-exact Context::flush (context.rs:50-92).

    * no colour: every glyph, then '\\n' (println!)
    * colour + webify: ``<span style="color:rgb(r,g,b)">c`` per cell, never closed
    * colour: crossterm 0.18 PrintStyledContent with fg = cell colour, bg = rgb(25,25,25):
      ``ESC[48;2;25;25;25m ESC[38;2;r;g;bm c ESC[0m`` (restated from crossterm 0.18's
      ``Display for StyledContent``: background, foreground, content, ResetColor; unpinned).
    Interactive mode additionally starts with ``ESC[1;1H`` (cursor::MoveTo(0,0)).
    """
    cells = np.asarray(cells, np.uint32)
    glyph = (cells & 0xFF).astype(np.uint8)
    out = bytearray()
    if not image:
        out += b"\x1b[1;1H"
    if not color:
        out += bytes(glyph) + b"\n"
        return bytes(out)
    r, g, b = (cells >> 8) & 0xFF, (cells >> 16) & 0xFF, (cells >> 24) & 0xFF
    if webify:
        for i in range(cells.size):
            out += b'<span style="color:rgb(%d,%d,%d)">' % (r[i], g[i], b[i])
            out.append(glyph[i])
    else:
        for i in range(cells.size):
            out += b"\x1b[48;2;25;25;25m\x1b[38;2;%d;%d;%dm" % (r[i], g[i], b[i])
            out.append(glyph[i])
            out += b"\x1b[0m"
    return bytes(out)
