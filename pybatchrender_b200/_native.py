"""ctypes binding of ``csrc/libpbr_b200.so`` (C ABI declared in ``include/pbr_b200.h``).

There is no CPU or PyTorch fallback behind this module: if the shared library is missing or the
tensors are not CUDA tensors the calls raise.  The library is built in-tree by
``__graft_entry__.build()`` / ``pybatchrender_b200.build.build_native()``.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PBR_B200_LIB: another build of the same library (A/B measurements of kernel variants on one box)
LIB_PATH = os.environ.get("PBR_B200_LIB") or os.path.join(_HERE, "csrc", "libpbr_b200.so")

PBR_MAX_NODES = 24
PBR_MESH_TWO_SIDED = 1
PBR_FRAME_FORCE_GENERAL = 1
PBR_FRAME_FORCE_FUSED = 2
PBR_FRAME_WRITE_MATS = 4
PBR_EOVERFLOW = -5
PBR_NODE_IN_BASE = 1


class NativeError(RuntimeError):
    pass


class _Channel(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("stride", ctypes.c_int32), ("constant", ctypes.c_float)]


class _PoseDesc(ctypes.Structure):
    _fields_ = [
        ("pos", _Channel * 3),
        ("hpr", _Channel * 3),
        ("scale", _Channel),
        ("out_mats", ctypes.c_void_p),
        ("n_instances", ctypes.c_int32),
    ]


class _NodeDesc(ctypes.Structure):
    _fields_ = [
        ("mesh", ctypes.c_void_p),
        ("mats", ctypes.c_void_p),
        ("cols", ctypes.c_void_p),
        ("instances_per_scene", ctypes.c_int32),
        ("shared", ctypes.c_int32),
        ("use_texture", ctypes.c_float),
        ("flags", ctypes.c_uint32),
        ("texture", ctypes.c_void_p),
        ("pose", ctypes.POINTER(_PoseDesc)),
    ]


class _FrameDesc(ctypes.Structure):
    _fields_ = [
        ("num_scenes", ctypes.c_int32),
        ("scene_begin", ctypes.c_int32),
        ("scene_count", ctypes.c_int32),
        ("tile_w", ctypes.c_int32),
        ("tile_h", ctypes.c_int32),
        ("channels", ctypes.c_int32),
        ("vp", ctypes.c_void_p),
        ("bg", ctypes.c_float * 4),
        ("ambient", ctypes.c_float * 3),
        ("dir_dir", ctypes.c_float * 3),
        ("dir_col", ctypes.c_float * 3),
        ("strength", ctypes.c_float),
        ("n_nodes", ctypes.c_int32),
        ("nodes", ctypes.POINTER(_NodeDesc)),
        ("out", ctypes.c_void_p),
        ("flags", ctypes.c_uint32),
        ("base", ctypes.c_void_p),
    ]


_EXPORTS = {
    "pbr_version": (ctypes.c_int, []),
    "pbr_last_error": (ctypes.c_char_p, []),
    "pbr_mesh_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                       ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32,
                                       ctypes.POINTER(ctypes.c_void_p)]),
    "pbr_mesh_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pbr_texture_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.POINTER(ctypes.c_void_p)]),
    "pbr_texture_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pbr_mesh_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32),
                                     ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "pbr_render": (ctypes.c_int, [ctypes.POINTER(_FrameDesc), ctypes.c_void_p]),
    "pbr_base_create": (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    "pbr_base_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pbr_base_render": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(_FrameDesc), ctypes.c_void_p]),
    "pbr_device_status": (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]),
    "pbr_pack_transforms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int32, ctypes.c_void_p]),
    "pbr_compose_transforms": (ctypes.c_int, [ctypes.POINTER(_PoseDesc), ctypes.c_int32, ctypes.c_void_p]),
    "pbr_device_status_nosync": (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    "pbr_kernel_launches": (ctypes.c_ulonglong, []),
}

_lib = None


def exported_symbols() -> list[str]:
    return sorted(_EXPORTS)


def load():
    """dlopen the library (raises NativeError with a build hint if it is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pbr_last_error()
        raise NativeError(f"{what} failed ({rc}): {msg.decode(errors='replace') if msg else ''}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)     # a plain C call: ~0.2 us instead of ~6 us


def _stream_ptr(device: torch.device) -> int:
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    return int(torch.cuda.current_stream(device).cuda_stream)


def _cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise NativeError(f"{what} must be a CUDA tensor: the B200 renderer has no CPU fallback")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise NativeError(f"{what} must be contiguous float32")
    return t


# Handles whose owner was garbage-collected while a CUDA graph was being captured: destroying them calls
# cudaFree, which is prohibited during a capture (it invalidates the capture in torch's default "global" error
# mode) -- and Python may collect an old renderer at any moment.  They are parked here and destroyed by the next
# create / close that runs outside a capture.
_deferred_frees: list = []


def _capturing() -> bool:
    try:
        return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    except Exception:
        return False


def _destroy(fn_name: str, handle) -> None:
    if _capturing():
        _deferred_frees.append((fn_name, handle))
        return
    lib = load()
    while _deferred_frees:
        name, h = _deferred_frees.pop()
        getattr(lib, name)(h)
    getattr(lib, fn_name)(handle)


class NativeMesh:
    """Device-resident static geometry (``pbr_mesh_t``)."""

    def __init__(self, pos, nrm, idx, device: torch.device, two_sided: bool = False, uv=None) -> None:
        import numpy as np
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        uv = None if uv is None else np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        if uv is not None and uv.shape[0] != pos.shape[0]:
            raise NativeError("uv must have one row per vertex")
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        handle = ctypes.c_void_p()
        rc = load().pbr_mesh_create(pos.ctypes.data, nrm.ctypes.data, None if uv is None else uv.ctypes.data,
                                    pos.shape[0], idx.ctypes.data,
                                    idx.shape[0], dev_index, PBR_MESH_TWO_SIDED if two_sided else 0,
                                    ctypes.byref(handle))
        _check(rc, "pbr_mesh_create")
        self.handle = handle
        self.n_tris = int(idx.shape[0])
        self.device_index = dev_index

    def close(self) -> None:
        if getattr(self, "handle", None):
            _destroy("pbr_mesh_destroy", self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NativeTexture:
    """Device-resident RGBA8 image (``pbr_texture_t``); ``rgba``: [h, w, 4] uint8, row 0 = v 0."""

    def __init__(self, rgba, device: torch.device) -> None:
        import numpy as np
        img = np.ascontiguousarray(rgba, dtype=np.uint8)
        if img.ndim != 3 or img.shape[2] != 4:
            raise NativeError("texture must be [h, w, 4] uint8")
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        handle = ctypes.c_void_p()
        _check(load().pbr_texture_create(img.ctypes.data, img.shape[1], img.shape[0], dev_index, ctypes.byref(handle)),
               "pbr_texture_create")
        self.handle = handle

    def close(self) -> None:
        if getattr(self, "handle", None):
            _destroy("pbr_texture_destroy", self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NativeBase:
    """Static layer handle (``pbr_base_t``): shared nodes under a uniform camera, rendered once."""

    def __init__(self, device: torch.device) -> None:
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        handle = ctypes.c_void_p()
        _check(load().pbr_base_create(dev_index, ctypes.byref(handle)), "pbr_base_create")
        self.handle = handle

    def close(self) -> None:
        if getattr(self, "handle", None):
            _destroy("pbr_base_destroy", self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def new_pose_struct(matbuf: torch.Tensor) -> "_PoseDesc":
    """A pbr_pose_desc for one node: identity pose, matrices materialise into ``matbuf``."""
    _cuda_f32(matbuf, "matbuf")
    st = _PoseDesc()
    for k in range(3):
        st.pos[k].ptr, st.pos[k].stride, st.pos[k].constant = None, 0, 0.0
        st.hpr[k].ptr, st.hpr[k].stride, st.hpr[k].constant = None, 0, 0.0
    st.scale.ptr, st.scale.stride, st.scale.constant = None, 0, 1.0
    st.out_mats = matbuf.data_ptr()
    st.n_instances = int(matbuf.shape[0])
    return st


class Native:
    """Thin object the renderer holds; every method enqueues work on torch's current stream."""

    def __init__(self) -> None:
        self.lib = load()

    def version(self) -> int:
        return int(self.lib.pbr_version())

    def kernel_launches(self) -> int:
        """Kernels enqueued or captured by the library in this process so far."""
        return int(self.lib.pbr_kernel_launches())

    def device_status_nosync(self, device_index: int) -> int:
        """The same bits from host-mapped memory, without synchronising (may lag behind frames in flight)."""
        v = ctypes.c_int32(0)
        _check(self.lib.pbr_device_status_nosync(int(device_index), ctypes.byref(v)), "pbr_device_status_nosync")
        return int(v.value)

    def device_status(self, device_index: int, clear: bool = True) -> int:
        """Sticky diagnostic bits written by the kernels (synchronises the device)."""
        v = ctypes.c_int32(0)
        _check(self.lib.pbr_device_status(int(device_index), ctypes.byref(v), 1 if clear else 0), "pbr_device_status")
        return int(v.value)

    def pack_transforms(self, transforms_b44, rot_b33, scale_b11, matbuf) -> None:
        n = int(matbuf.shape[0])
        for t, w in ((transforms_b44, "transforms_b44"), (rot_b33, "rot3_b33"), (scale_b11, "scale_b11"),
                     (matbuf, "matbuf")):
            _cuda_f32(t, w)
            if t.shape[0] != n:
                raise NativeError(f"{w} has {t.shape[0]} rows, matbuf has {n}")
        with torch.cuda.device(matbuf.device):
            rc = self.lib.pbr_pack_transforms(transforms_b44.data_ptr(), rot_b33.data_ptr(), scale_b11.data_ptr(),
                                              matbuf.data_ptr(), n, _stream_ptr(matbuf.device))
        _check(rc, "pbr_pack_transforms")

    @staticmethod
    def fill_pose(dst: "_PoseDesc", p: dict) -> None:
        """p: {pos: (c,c,c), hpr: (c,c,c), scale: c, out: tensor[B,16]}, c = float | 1-D float32 CUDA view with one
        element per instance (the caller keeps the tensors alive)."""
        out = _cuda_f32(p["out"], "pose out")
        n = int(out.shape[0])

        def chan(d, v):
            if isinstance(v, torch.Tensor):
                if not v.is_cuda or v.dtype != torch.float32 or v.dim() != 1:
                    raise NativeError("pose channel tensors must be 1-D float32 CUDA views")
                if v.shape[0] != n:
                    raise NativeError(f"pose channel has {v.shape[0]} elements, node has {n} instances")
                d.ptr, d.stride, d.constant = v.data_ptr(), int(v.stride(0)), 0.0
            else:
                d.ptr, d.stride, d.constant = None, 0, float(v)

        for k in range(3):
            chan(dst.pos[k], p["pos"][k])
            chan(dst.hpr[k], p["hpr"][k])
        chan(dst.scale, p.get("scale", 1.0))
        dst.out_mats = out.data_ptr()
        dst.n_instances = n

    def compose_structs(self, structs: list, device: torch.device) -> None:
        """pbr_compose_transforms over ready-made pbr_pose_desc structures (``new_pose_struct``)."""
        arr = (_PoseDesc * len(structs))(*structs)
        with torch.cuda.device(device):
            rc = self.lib.pbr_compose_transforms(arr, len(structs), _stream_ptr(device))
        _check(rc, "pbr_compose_transforms")

    def compose(self, poses: list[dict], device: torch.device) -> None:
        """Write the matrices of several poses (see ``fill_pose``) in one launch."""
        arr = (_PoseDesc * len(poses))()
        for i, p in enumerate(poses):
            self.fill_pose(arr[i], p)
        with torch.cuda.device(device):
            rc = self.lib.pbr_compose_transforms(arr, len(poses), _stream_ptr(device))
        _check(rc, "pbr_compose_transforms")

    def _frame(self, *, num_scenes, tile_w, tile_h, channels, vp, nodes, out, bg, ambient, dir_dir, dir_col,
               strength, scene_begin=0, scene_count=None, flags=0, base=None):
        """nodes: list of (NativeMesh, matbuf, colbuf, instances_per_scene, shared[, in_base[, use_texture,
        NativeTexture | None[, pose dict | pbr_pose_desc structure | None]]])."""
        _cuda_f32(vp, "viewbuf")
        nd = (_NodeDesc * max(1, len(nodes)))()
        poses = []                     # pbr_pose_desc structures the node array points at (kept alive with it)
        for i, item in enumerate(nodes):
            mesh, mats, cols, inst, shared = item[:5]
            _cuda_f32(mats, "matbuf")
            _cuda_f32(cols, "colbuf")
            nd[i].mesh = mesh.handle
            nd[i].mats = mats.data_ptr()
            nd[i].cols = cols.data_ptr()
            nd[i].instances_per_scene = int(inst)
            nd[i].shared = 1 if shared else 0
            nd[i].use_texture = float(item[6]) if len(item) > 6 else 0.0
            nd[i].texture = item[7].handle if len(item) > 7 and item[7] is not None else None
            nd[i].flags = PBR_NODE_IN_BASE if (len(item) > 5 and item[5]) else 0
            if len(item) > 8 and item[8] is not None:
                st = item[8]
                if isinstance(st, dict):
                    d, st = st, _PoseDesc()
                    self.fill_pose(st, d)
                poses.append(st)
                nd[i].pose = ctypes.pointer(st)
        f = _FrameDesc()
        f.num_scenes = int(num_scenes)
        f.scene_begin = int(scene_begin)
        f.scene_count = int(num_scenes - scene_begin if scene_count is None else scene_count)
        f.tile_w, f.tile_h, f.channels = int(tile_w), int(tile_h), int(channels)
        f.vp = vp.data_ptr()
        f.bg = (ctypes.c_float * 4)(*[float(x) for x in bg])
        f.ambient = (ctypes.c_float * 3)(*[float(x) for x in ambient])
        f.dir_dir = (ctypes.c_float * 3)(*[float(x) for x in dir_dir])
        f.dir_col = (ctypes.c_float * 3)(*[float(x) for x in dir_col])
        f.strength = float(strength)
        f.n_nodes = len(nodes)
        f.nodes = ctypes.cast(nd, ctypes.POINTER(_NodeDesc))
        f.out = out.data_ptr() if out is not None else None
        f.flags = int(flags)
        f.base = base.handle if base is not None else None
        return f, (nd, poses)

    def prepare(self, *, out, **kw):
        """Build the native description of a frame once; ``render_cached`` enqueues it any number of times (with
        another output buffer / scene window).  Valid while the tensors it points at are alive and the scene's
        structure is unchanged -- the renderer keys it on its version counters."""
        if not out.is_cuda or out.dtype != torch.uint8 or not out.is_contiguous():
            raise NativeError("out must be a contiguous uint8 CUDA tensor")
        f, keep = self._frame(out=out, **kw)
        dev = out.device
        return [f, ctypes.byref(f), keep, dev, dev.index if dev.index is not None else torch.cuda.current_device(),
                int(kw["num_scenes"])]

    def render_cached(self, prepared, out, scene_begin: int = 0, scene_count=None) -> None:
        f, fref, _keep, dev, dev_index, n = prepared
        if not out.is_cuda or out.dtype != torch.uint8 or not out.is_contiguous():
            raise NativeError("out must be a contiguous uint8 CUDA tensor")
        f.out = out.data_ptr()
        f.scene_begin = scene_begin
        f.scene_count = n - scene_begin if scene_count is None else scene_count
        stream = _raw_stream(dev_index) if _raw_stream is not None else torch.cuda.current_stream(dev).cuda_stream
        if torch.cuda.current_device() == dev_index:
            rc = self.lib.pbr_render(fref, stream)
        else:
            with torch.cuda.device(dev):
                rc = self.lib.pbr_render(fref, stream)
        if rc != 0:
            _check(rc, "pbr_render")

    def render(self, *, out, **kw) -> None:
        if not out.is_cuda or out.dtype != torch.uint8 or not out.is_contiguous():
            raise NativeError("out must be a contiguous uint8 CUDA tensor")
        f, keep = self._frame(out=out, **kw)
        with torch.cuda.device(out.device):
            rc = self.lib.pbr_render(ctypes.byref(f), _stream_ptr(out.device))
        _check(rc, "pbr_render")
        del keep

    def base_render(self, base: NativeBase, *, vp, **kw) -> None:
        """Render the static layer from the nodes flagged in_base (vp row ``scene_begin`` is used)."""
        f, keep = self._frame(out=None, vp=vp, **kw)
        with torch.cuda.device(vp.device):
            rc = self.lib.pbr_base_render(base.handle, ctypes.byref(f), _stream_ptr(vp.device))
        _check(rc, "pbr_base_render")
        del keep
