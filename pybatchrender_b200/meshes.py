"""Mesh sources and the model "bake" the reference performs when a node is created.

Reference behaviour being mirrored (``pybatchrender/renderer/node.py:61-72, 214-268, 313-341``):

1. ``loader.loadModel(path)``
2. ``setScale`` -- ``model_scale_units="relative"`` multiplies, ``"absolute"`` rescales the tight
   bounds to the requested size (uniform: longest axis; per-axis otherwise)
3. ``setHpr(model_hpr)`` (Panda3D convention, degrees: H about +Z, P about +X, R about +Y)
4. ``flattenStrong`` -- the transform is baked into vertices and normals
5. ``pivot_to_rel(rel)`` -- translate so the bounds-relative point ``min + (max-min)*rel`` becomes
   the origin, bake again.

``models/box`` is a Panda3D built-in that is not vendored in the reference tree; the notebook
goldens pin it as the unit cube ``[0,1]^3`` with per-face normals, which is
what :func:`box` builds.  ``models/smiley`` (also a built-in) is replaced by a UV sphere of the
same radius (parity unpinned: no reference output exists for it).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np


@dataclass
class MeshData:
    """Indexed triangle mesh in object space.  pos/nrm: [V,3] f32, idx: [T,3] u32."""
    pos: np.ndarray
    nrm: np.ndarray
    idx: np.ndarray
    uv: np.ndarray | None = None
    two_sided: bool = False

    def copy(self) -> "MeshData":
        return MeshData(self.pos.copy(), self.nrm.copy(), self.idx.copy(),
                        None if self.uv is None else self.uv.copy(), self.two_sided)

    @property
    def n_tris(self) -> int:
        return int(self.idx.shape[0])

    def tight_bounds(self) -> tuple[np.ndarray, np.ndarray]:
        return self.pos.min(axis=0), self.pos.max(axis=0)


def box() -> MeshData:
    """Unit cube [0,1]^3, 24 vertices (4 per face, per-face normals), 12 CCW triangles."""
    faces = [
        # normal, origin corner, u axis, v axis  (u x v = normal -> CCW seen from outside)
        ((+1, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)),
        ((-1, 0, 0), (0, 0, 0), (0, 0, 1), (0, 1, 0)),
        ((0, +1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 0)),
        ((0, -1, 0), (0, 0, 0), (1, 0, 0), (0, 0, 1)),
        ((0, 0, +1), (0, 0, 1), (1, 0, 0), (0, 1, 0)),
        ((0, 0, -1), (0, 0, 0), (0, 1, 0), (1, 0, 0)),
    ]
    pos, nrm, uv, idx = [], [], [], []
    for n, o, u, v in faces:
        o, u, v = np.array(o, np.float32), np.array(u, np.float32), np.array(v, np.float32)
        base = len(pos)
        for (a, b) in ((0, 0), (1, 0), (1, 1), (0, 1)):
            pos.append(o + a * u + b * v)
            nrm.append(n)
            uv.append((a, b))
        idx.append((base, base + 1, base + 2))
        idx.append((base, base + 2, base + 3))
    return MeshData(np.array(pos, np.float32), np.array(nrm, np.float32), np.array(idx, np.uint32),
                    np.array(uv, np.float32))


def uv_sphere(radius: float = 1.0, segments: int = 16, rings: int = 12) -> MeshData:
    """Smooth-shaded UV sphere centred on the origin (stand-in for Panda3D's ``models/smiley``)."""
    pos, nrm, uv, idx = [], [], [], []
    for r in range(rings + 1):
        phi = math.pi * r / rings
        for s in range(segments + 1):
            th = 2.0 * math.pi * s / segments
            n = (math.sin(phi) * math.cos(th), math.sin(phi) * math.sin(th), math.cos(phi))
            nrm.append(n)
            pos.append(tuple(radius * c for c in n))
            uv.append((s / segments, 1.0 - r / rings))
    row = segments + 1
    for r in range(rings):
        for s in range(segments):
            a, b = r * row + s, r * row + s + 1
            c, d = (r + 1) * row + s, (r + 1) * row + s + 1
            if r != 0:
                idx.append((a, c, b))
            if r != rings - 1:
                idx.append((b, c, d))
    return MeshData(np.array(pos, np.float32), np.array(nrm, np.float32), np.array(idx, np.uint32),
                    np.array(uv, np.float32))


def cone(segments: int = 32) -> MeshData:
    """Cone in the frame of the reference's ``models/cone.egg``: base circle of radius 1 in the
    plane y = -1 (flat cap, normal -y), apex at (0, 1, 0), smooth side normals.  The apex gets one
    vertex per side triangle carrying the side normal at the triangle's mid angle."""
    pos, nrm, tris = [], [], []
    slope = 1.0 / math.sqrt(5.0)                      # side normal = (2 cos, 1, 2 sin) / sqrt(5)
    ring = [(math.sin(2.0 * math.pi * k / segments), -math.cos(2.0 * math.pi * k / segments)) for k in range(segments)]
    for x, z in ring:                                 # cap ring: vertices 0 .. segments-1
        pos.append((x, -1.0, z))
        nrm.append((0.0, -1.0, 0.0))
    for x, z in ring:                                 # side ring: segments .. 2*segments-1
        pos.append((x, -1.0, z))
        nrm.append((2.0 * slope * x, slope, 2.0 * slope * z))
    for k in range(segments):                         # apex copies: 2*segments .. 3*segments-1
        a = 2.0 * math.pi * (k + 0.5) / segments
        pos.append((0.0, 1.0, 0.0))
        nrm.append((2.0 * slope * math.sin(a), slope, -2.0 * slope * math.cos(a)))
    for k in range(1, segments - 1):                  # cap fan (what a loader makes of the n-gon)
        tris.append((0, k, k + 1))
    for k in range(segments):
        tris.append((segments + k, 2 * segments + k, segments + (k + 1) % segments))
    return MeshData(np.array(pos, np.float32), np.array(nrm, np.float32), np.array(tris, np.uint32))


def cylinder(radius: float = 50.0, half_height: float = 100.0, segments: int = 32, stacks: int = 4) -> MeshData:
    """Capped cylinder about the +Z axis (the shape of the reference's ``models/cylinder/scene.gltf``
    once loaded into a Z-up world: radius 50, z in [-100, 100], 32 segments x 4 stacks, flat caps,
    smooth side): 2 + segments * (stacks + 3) vertices, segments * (2 * stacks + 2) triangles."""
    pos, nrm, tris = [], [], []
    pos += [(0.0, 0.0, -half_height), (0.0, 0.0, half_height)]
    nrm += [(0.0, 0.0, -1.0), (0.0, 0.0, 1.0)]
    col = stacks + 3                                  # per segment: bottom cap, stacks+1 side, top cap
    for k in range(segments):
        a = 2.0 * math.pi * k / segments
        c, s_ = math.cos(a), math.sin(a)
        pos.append((radius * c, radius * s_, -half_height)); nrm.append((0.0, 0.0, -1.0))
        for j in range(stacks + 1):
            pos.append((radius * c, radius * s_, -half_height + 2.0 * half_height * j / stacks)); nrm.append((c, s_, 0.0))
        pos.append((radius * c, radius * s_, half_height)); nrm.append((0.0, 0.0, 1.0))
    for k in range(segments):
        a0, a1 = 2 + k * col, 2 + ((k + 1) % segments) * col
        tris.append((0, a1, a0))                                            # bottom cap (faces -z)
        for j in range(stacks):
            p00, p01, p10, p11 = a0 + 1 + j, a0 + 2 + j, a1 + 1 + j, a1 + 2 + j
            tris.append((p10, p11, p01))
            tris.append((p10, p01, p00))
        tris.append((1, a0 + col - 1, a1 + col - 1))                        # top cap (faces +z)
    m = MeshData(np.array(pos, np.float32), np.array(nrm, np.float32), np.array(tris, np.uint32))
    m.two_sided = True                                # the asset's material is doubleSided
    return m


MODELS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")

_BUILTINS = {
    "models/box": box,
    "box": box,
    "models/smiley": uv_sphere,
    "smiley": uv_sphere,
    "models/sphere": uv_sphere,
}

_registry: dict[str, MeshData] = {}


def register_mesh(name: str, mesh: MeshData) -> None:
    """Make ``add_node(name, ...)`` resolve to ``mesh`` (user-supplied geometry)."""
    _registry[name] = mesh


def load_mesh(model_path) -> MeshData:
    """Resolve what the reference hands to ``loader.loadModel`` (node.py:62)."""
    if isinstance(model_path, MeshData):
        return model_path.copy()
    key = str(model_path)
    if key in _registry:
        return _registry[key].copy()
    stem = key[:-4] if key.endswith((".egg", ".bam")) and key[:-4] in _BUILTINS else key
    if stem in _BUILTINS:
        return _BUILTINS[stem]()
    from . import mesh_io
    if os.path.isfile(key):
        return mesh_io.load_file(key)
    # "models/<file>" resolves into the package's models directory the way the reference's Steering
    # renderer resolves it (envs/steering/renderer.py:90-107)
    if key.startswith("models/") and os.path.isfile(os.path.join(MODELS_DIR, key[7:])):
        return mesh_io.load_file(os.path.join(MODELS_DIR, key[7:]))
    raise FileNotFoundError(
        f"model {model_path!r}: not a built-in ({sorted(set(_BUILTINS))}), not registered and not a file")


# ---------------------------------------------------------------------------- bake
def _panda_hpr_matrix(hpr_deg) -> np.ndarray:
    """Panda3D ``setHpr`` rotation (degrees): roll about Y, then pitch about X, then heading about Z."""
    h, p, r = (math.radians(float(a)) for a in hpr_deg)
    ch, sh, cp, sp, cr, sr = math.cos(h), math.sin(h), math.cos(p), math.sin(p), math.cos(r), math.sin(r)
    rz = np.array([[ch, -sh, 0], [sh, ch, 0], [0, 0, 1]], np.float64)
    rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]], np.float64)
    ry = np.array([[cr, 0, sr], [0, 1, 0], [-sr, 0, cr]], np.float64)
    return rz @ rx @ ry


def coerce_scale(scale):
    """-> ((sx,sy,sz), was_uniform) or None   (node.py:180-200)."""
    if scale is None:
        return None
    if isinstance(scale, (int, float)):
        v = float(scale)
        return (v, v, v), True
    seq = tuple(float(v) for v in scale)
    if len(seq) == 0:
        raise ValueError("model_scale sequence cannot be empty")
    if len(seq) == 1:
        return (seq[0],) * 3, True
    if len(seq) != 3:
        raise ValueError("model_scale sequence must have length 1 or 3")
    return seq, False


def coerce_hpr(hpr):
    if hpr is None:
        return None
    try:
        seq = tuple(float(v) for v in hpr)
    except TypeError:
        raise TypeError("model_hpr must be a sequence of length 3 or None")
    if len(seq) != 3:
        raise ValueError("model_hpr sequence must have length 3")
    return seq


def apply_linear(mesh: MeshData, lin: np.ndarray, offset=None) -> None:
    """Bake ``v -> lin @ v + offset`` into the mesh (what ``flattenStrong`` does)."""
    lin = np.asarray(lin, np.float64)
    pos = mesh.pos.astype(np.float64) @ lin.T
    if offset is not None:
        pos = pos + np.asarray(offset, np.float64)
    nmat = np.linalg.inv(lin).T
    nrm = mesh.nrm.astype(np.float64) @ nmat.T
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = nrm / np.where(ln > 0, ln, 1.0)
    mesh.pos = pos.astype(np.float32)
    mesh.nrm = nrm.astype(np.float32)
    if np.linalg.det(lin) < 0:          # mirrored: keep faces front-facing
        mesh.idx = mesh.idx[:, [0, 2, 1]].copy()


def bake(mesh: MeshData, model_scale=None, model_hpr=None, model_scale_units="relative",
         pivot_rel=None) -> MeshData:
    """Steps 2-5 of the module docstring, in the reference's order."""
    sc = coerce_scale(model_scale)
    lin = np.eye(3)
    if sc is not None:
        values, uniform = sc
        if model_scale_units == "relative":
            factors = values
        elif model_scale_units == "absolute":
            lo, hi = mesh.tight_bounds()
            dims = (hi - lo).astype(np.float64)
            longest = max(float(dims.max()), 1e-8)
            if uniform:
                factors = (values[0] / longest,) * 3
            else:
                factors = []
                for target, src in zip(values, dims):
                    den = src if abs(src) >= 1e-8 else longest
                    if abs(den) < 1e-8:
                        den = 1.0
                    factors.append(target / den)
        else:
            raise ValueError(f"Unknown model_scale_units '{model_scale_units}'")
        lin = np.diag([float(f) for f in factors])
    hpr = coerce_hpr(model_hpr)
    if hpr is not None:
        lin = _panda_hpr_matrix(hpr) @ lin      # Panda composes scale first, then rotation
    if sc is not None or hpr is not None:
        apply_linear(mesh, lin)
    if pivot_rel is not None:
        lo, hi = mesh.tight_bounds()
        p = lo.astype(np.float64) + (hi - lo).astype(np.float64) * np.asarray(pivot_rel, np.float64)
        mesh.pos = (mesh.pos.astype(np.float64) - p).astype(np.float32)
    return mesh
