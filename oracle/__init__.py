"""ctypes binding of the CPU oracle (``oracle/pbr_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
``pybatchrender_b200`` never does (tests/test_no_oracle_in_product.py enforces it).

The oracle restates the reference's pixel path (``pybatchrender/shaders/basic.vert``,
``basic.frag`` and the GL fixed-function rules) on the CPU and is
pinned on the notebook goldens (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpbr_oracle.so")
_lib = None

MESH_TWO_SIDED = 1


class _Mesh(ctypes.Structure):
    _fields_ = [
        ("pos", ctypes.c_void_p),
        ("nrm", ctypes.c_void_p),
        ("idx", ctypes.c_void_p),
        ("n_verts", ctypes.c_int),
        ("n_tris", ctypes.c_int),
        ("flags", ctypes.c_uint32),
        ("uv", ctypes.c_void_p),
    ]


class _Node(ctypes.Structure):
    _fields_ = [
        ("mesh", _Mesh),
        ("mats", ctypes.c_void_p),
        ("cols", ctypes.c_void_p),
        ("instances_per_scene", ctypes.c_int),
        ("shared", ctypes.c_int),
        ("use_texture", ctypes.c_float),
        ("tex", ctypes.c_void_p),
        ("tex_w", ctypes.c_int),
        ("tex_h", ctypes.c_int),
    ]


class _Frame(ctypes.Structure):
    _fields_ = [
        ("num_scenes", ctypes.c_int),
        ("scene_begin", ctypes.c_int),
        ("scene_count", ctypes.c_int),
        ("tile_w", ctypes.c_int),
        ("tile_h", ctypes.c_int),
        ("channels", ctypes.c_int),
        ("vp", ctypes.c_void_p),
        ("bg", ctypes.c_float * 4),
        ("ambient", ctypes.c_float * 3),
        ("dir_dir", ctypes.c_float * 3),
        ("dir_col", ctypes.c_float * 3),
        ("strength", ctypes.c_float),
        ("n_nodes", ctypes.c_int),
        ("nodes", ctypes.POINTER(_Node)),
        ("out", ctypes.c_void_p),
        ("n_threads", ctypes.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile ``libpbr_oracle.so`` with the committed Makefile (gcc, no GPU needed)."""
    src = os.path.join(_HERE, "pbr_oracle.c")
    hdr = os.path.join(_HERE, "pbr_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_render.argtypes = [ctypes.POINTER(_Frame)]
        _lib.orc_render.restype = ctypes.c_int
        _lib.orc_render_scene_debug.argtypes = [
            ctypes.POINTER(_Frame), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.orc_render_scene_debug.restype = ctypes.c_int
        _lib.orc_version.restype = ctypes.c_int
    return _lib


def _f32(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


@dataclass
class OracleNode:
    """One ``PBRNode`` worth of shader inputs (reference ``node.py:85-91``)."""
    pos: np.ndarray          # [V,3] baked object-space vertices
    nrm: np.ndarray          # [V,3]
    idx: np.ndarray          # [T,3] uint32
    mats: np.ndarray         # [B,16] column-packed (== matbuf)
    cols: np.ndarray         # [B,4]                (== colbuf)
    instances_per_scene: int = 1
    shared: bool = False
    flags: int = 0
    uv: np.ndarray | None = None        # [V,2]
    texture: np.ndarray | None = None   # [h,w,4] uint8, row 0 = v 0
    use_texture: float = 0.0


@dataclass
class OracleFrame:
    """Everything one ``renderer.step(return_pixels=True)`` consumes below the host classes."""
    num_scenes: int
    tile_w: int
    tile_h: int
    vp: np.ndarray                         # [K,16] column-packed (== viewbuf)
    nodes: list = field(default_factory=list)
    channels: int = 3
    bg: tuple = (0.0, 0.0, 0.0, 1.0)
    ambient: tuple = (0.2, 0.2, 0.25)
    dir_dir: tuple = (0.4, -0.6, -0.7)
    dir_col: tuple = (1.0, 1.0, 1.0)
    strength: float = 1.0


def _pack(frame: OracleFrame, out: np.ndarray, scene_begin: int, scene_count: int, n_threads: int):
    keep = []
    K = int(frame.num_scenes)
    vp = _f32(frame.vp, (K, 16))
    keep.append(vp)
    nodes = (_Node * max(1, len(frame.nodes)))()
    for i, n in enumerate(frame.nodes):
        pos, nrm = _f32(n.pos, (-1, 3)), _f32(n.nrm, (-1, 3))
        idx = np.ascontiguousarray(np.asarray(n.idx, dtype=np.uint32).reshape(-1, 3))
        B = n.instances_per_scene if n.shared else K * n.instances_per_scene
        mats, cols = _f32(n.mats, (-1, 16)), _f32(n.cols, (-1, 4))
        if mats.shape[0] < B or cols.shape[0] < B:
            raise ValueError(f"node {i}: need {B} instance rows, got {mats.shape[0]} / {cols.shape[0]}")
        if idx.size and int(idx.max()) >= pos.shape[0]:
            raise ValueError("index out of range")
        keep += [pos, nrm, idx, mats, cols]
        uv = None if n.uv is None else _f32(n.uv, (-1, 2))
        tex = None if n.texture is None else np.ascontiguousarray(np.asarray(n.texture, dtype=np.uint8))
        if tex is not None and (tex.ndim != 3 or tex.shape[2] != 4):
            raise ValueError(f"node {i}: texture must be [h,w,4] uint8")
        keep += [uv, tex]
        nodes[i].mesh = _Mesh(pos.ctypes.data, nrm.ctypes.data, idx.ctypes.data,
                              pos.shape[0], idx.shape[0], int(n.flags), None if uv is None else uv.ctypes.data)
        nodes[i].use_texture = float(n.use_texture)
        nodes[i].tex = None if tex is None else tex.ctypes.data
        nodes[i].tex_w = 0 if tex is None else int(tex.shape[1])
        nodes[i].tex_h = 0 if tex is None else int(tex.shape[0])
        nodes[i].mats = mats.ctypes.data
        nodes[i].cols = cols.ctypes.data
        nodes[i].instances_per_scene = int(n.instances_per_scene)
        nodes[i].shared = 1 if n.shared else 0
    f = _Frame()
    f.num_scenes, f.scene_begin, f.scene_count = K, int(scene_begin), int(scene_count)
    f.tile_w, f.tile_h, f.channels = int(frame.tile_w), int(frame.tile_h), int(frame.channels)
    f.vp = vp.ctypes.data
    f.bg = (ctypes.c_float * 4)(*[float(x) for x in frame.bg])
    f.ambient = (ctypes.c_float * 3)(*[float(x) for x in frame.ambient])
    f.dir_dir = (ctypes.c_float * 3)(*[float(x) for x in frame.dir_dir])
    f.dir_col = (ctypes.c_float * 3)(*[float(x) for x in frame.dir_col])
    f.strength = float(frame.strength)
    f.n_nodes = len(frame.nodes)
    f.nodes = ctypes.cast(nodes, ctypes.POINTER(_Node))
    f.out = out.ctypes.data
    f.n_threads = int(n_threads)
    keep.append(nodes)
    return f, keep


class PackedFrame:
    """A frame marshalled once and rendered many times (bench.py's CPU legs): the ctypes structures keep
    pointing at the numpy arrays handed in, so arrays that alias live buffers (e.g. ``tensor.numpy()`` of a
    renderer's CPU matrix buffers) are re-read on every ``render()`` without any per-frame Python work."""

    def __init__(self, frame: OracleFrame, out: np.ndarray, n_threads: int = 1, scene_begin: int = 0,
                 scene_count: int | None = None) -> None:
        K = int(frame.num_scenes)
        assert out.dtype == np.uint8 and out.flags.c_contiguous
        self.out = out
        self._f, self._keep = _pack(frame, out, scene_begin, K - scene_begin if scene_count is None else scene_count,
                                    n_threads)
        self._lib = lib()

    def render(self) -> np.ndarray:
        rc = self._lib.orc_render(ctypes.byref(self._f))
        if rc != 0:
            raise ValueError(f"orc_render failed: {rc}")
        return self.out


def render(frame: OracleFrame, n_threads: int = 1, scene_begin: int = 0, scene_count: int | None = None,
           out: np.ndarray | None = None) -> np.ndarray:
    """Render scenes [scene_begin, scene_begin+scene_count) -> uint8 [K,C,H,W] (other rows untouched)."""
    K = int(frame.num_scenes)
    if scene_count is None:
        scene_count = K - scene_begin
    if out is None:
        out = np.zeros((K, frame.channels, frame.tile_h, frame.tile_w), dtype=np.uint8)
    assert out.dtype == np.uint8 and out.flags.c_contiguous
    f, keep = _pack(frame, out, scene_begin, scene_count, n_threads)
    rc = lib().orc_render(ctypes.byref(f))
    if rc != 0:
        raise ValueError(f"orc_render failed: {rc}")
    del keep
    return out


def render_scene_debug(frame: OracleFrame, scene: int):
    """-> (pixels [C,H,W] u8, depth [H,W] f32, prim [H,W] u32) for one scene."""
    H, W, C = frame.tile_h, frame.tile_w, frame.channels
    px = np.zeros((C, H, W), np.uint8)
    depth = np.zeros((H, W), np.float32)
    prim = np.zeros((H, W), np.uint32)
    dummy = np.zeros(1, np.uint8)
    f, keep = _pack(frame, dummy, 0, 0, 1)
    rc = lib().orc_render_scene_debug(ctypes.byref(f), int(scene), px.ctypes.data, depth.ctypes.data,
                                      prim.ctypes.data)
    if rc != 0:
        raise ValueError(f"orc_render_scene_debug failed: {rc}")
    del keep
    return px, depth, prim
