"""Model files -> :class:`MeshData`: the two on-disk formats the reference hands to Panda3D's loader.

The reference calls ``loader.loadModel(path)`` (``pybatchrender/renderer/node.py:62``) on
``models/cone.egg`` (Panda3D's text format, written by ``obj2egg``) and
``models/cylinder/scene.gltf`` (glTF 2.0 + ``scene.bin``) -- ``envs/steering/config.py:56-59``.
Panda3D is not a dependency here, so this module reads the subset of both formats those assets (and
typical exports) use and returns geometry in Panda3D's frame (Z up, Y forward):

``.egg``
    ``<CoordinateSystem>``, ``<VertexPool>`` / ``<Vertex>`` with ``<Normal>`` / ``<UV>``,
    ``<Polygon>`` with ``<VertexRef>`` / ``<Normal>`` / ``<BFace>``, ``<Group>`` nesting and
    ``<Instance>`` transforms.  Polygons with more than three corners are fan-triangulated from
    their first corner (exact for the convex caps ``obj2egg`` writes).
``.gltf`` / ``.glb``
    scenes / nodes (``matrix`` or TRS), meshes / primitives (triangles, strips, fans), accessors
    through buffer views (``byteStride``, ``byteOffset``), external / base64 / GLB buffers,
    ``doubleSided`` materials.  Positions are moved from glTF's Y-up into Z-up the way Panda3D's
    importers do: ``(x, y, z) -> (x, -z, y)``.

Vertices without normals get the geometric normal of their face (flat shading).  ``write_egg`` /
``write_gltf`` are the inverse, used by ``tools/make_models.py`` to produce the bundled assets and
by the tests for round trips.
"""
from __future__ import annotations

import base64
import json
import math
import os
import re
import struct

import numpy as np

from .meshes import MeshData

__all__ = ["load_file", "load_egg", "load_gltf", "write_egg", "write_gltf"]


def load_file(path) -> MeshData:
    p = os.fspath(path)
    ext = os.path.splitext(p)[1].lower()
    if ext == ".egg":
        return load_egg(p)
    if ext in (".gltf", ".glb"):
        return load_gltf(p)
    raise ValueError(f"{p}: unsupported model format {ext!r} (supported: .egg, .gltf, .glb)")


# ------------------------------------------------------------------------------------------ helpers
def _face_normal(pts: np.ndarray) -> np.ndarray:
    """Newell normal of a planar polygon (robust for any corner count)."""
    nxt = np.roll(pts, -1, axis=0)
    n = np.array([
        np.sum((pts[:, 1] - nxt[:, 1]) * (pts[:, 2] + nxt[:, 2])),
        np.sum((pts[:, 2] - nxt[:, 2]) * (pts[:, 0] + nxt[:, 0])),
        np.sum((pts[:, 0] - nxt[:, 0]) * (pts[:, 1] + nxt[:, 1])),
    ], dtype=np.float64)
    ln = float(np.linalg.norm(n))
    return n / ln if ln > 0 else np.array([0.0, 0.0, 1.0])


class _Builder:
    """Accumulates corners, sharing a vertex when position, normal and uv are identical."""

    def __init__(self):
        self.pos, self.nrm, self.uv, self.tris = [], [], [], []
        self._seen: dict[bytes, int] = {}
        self.has_uv = False

    def corner(self, p, n, uv) -> int:
        rec = np.empty(8, dtype=np.float32)
        rec[0:3] = p
        rec[3:6] = n
        rec[6:8] = (0.0, 0.0) if uv is None else uv
        key = rec.tobytes()
        i = self._seen.get(key)
        if i is None:
            i = len(self.pos)
            self._seen[key] = i
            self.pos.append(rec[0:3].copy())
            self.nrm.append(rec[3:6].copy())
            self.uv.append(rec[6:8].copy())
        if uv is not None:
            self.has_uv = True
        return i

    def mesh(self, two_sided: bool) -> MeshData:
        if not self.tris:
            raise ValueError("model contains no triangles")
        return MeshData(np.asarray(self.pos, np.float32).reshape(-1, 3), np.asarray(self.nrm, np.float32).reshape(-1, 3),
                        np.asarray(self.tris, np.uint32).reshape(-1, 3),
                        np.asarray(self.uv, np.float32).reshape(-1, 2) if self.has_uv else None, two_sided)


# ---------------------------------------------------------------------------------------------- egg
_EGG_TOKEN = re.compile(r'"((?:[^"\\]|\\.)*)"|(<[^<>\s]+>)|([{}])|([^\s{}"]+)')


class _Egg:
    __slots__ = ("tag", "name", "values", "children")

    def __init__(self, tag, name):
        self.tag, self.name, self.values, self.children = tag, name, [], []

    def child(self, tag):
        for c in self.children:
            if c.tag == tag:
                return c
        return None

    def floats(self):
        return [float(v) for v in self.values]


def _egg_parse(text: str) -> _Egg:
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    toks = []
    for m in _EGG_TOKEN.finditer(text):
        if m.group(1) is not None:
            toks.append(("s", m.group(1)))
        elif m.group(2) is not None:
            toks.append(("t", m.group(2)[1:-1].lower()))
        elif m.group(3) is not None:
            toks.append((m.group(3), m.group(3)))
        else:
            toks.append(("w", m.group(4)))
    root = _Egg("root", "")
    stack = [root]
    i, n = 0, len(toks)
    while i < n:
        kind, val = toks[i]
        if kind == "t":
            name = []
            i += 1
            while i < n and toks[i][0] in ("w", "s"):
                name.append(toks[i][1])
                i += 1
            if i >= n or toks[i][0] != "{":
                raise ValueError(f"egg: <{val}> without a body")
            node = _Egg(val, " ".join(name))
            stack[-1].children.append(node)
            stack.append(node)
        elif kind == "}":
            if len(stack) == 1:
                raise ValueError("egg: unbalanced '}'")
            stack.pop()
        elif kind == "{":
            raise ValueError("egg: unexpected '{'")
        else:
            stack[-1].values.append(val)
        i += 1
    if len(stack) != 1:
        raise ValueError("egg: unbalanced '{'")
    return root


def _egg_transform(node: _Egg) -> np.ndarray:
    """Net matrix of a <Transform> body, row-vector convention (v' = v @ M), components in order."""
    m = np.eye(4)
    for c in node.children:
        v = c.floats()
        t = np.eye(4)
        if c.tag == "matrix4" and len(v) == 16:
            t = np.array(v).reshape(4, 4)
        elif c.tag == "matrix3" and len(v) == 9:
            a = np.array(v).reshape(3, 3)
            t[0:2, 0:2] = a[0:2, 0:2]
            t[3, 0:2] = a[2, 0:2]
        elif c.tag == "translate":
            t[3, 0:len(v)] = v
        elif c.tag == "scale":
            s = v * 3 if len(v) == 1 else v
            t[0, 0], t[1, 1], t[2, 2] = s[0], s[1], s[2]
        elif c.tag in ("rotx", "roty", "rotz", "rotate"):
            axis = {"rotx": (1, 0, 0), "roty": (0, 1, 0), "rotz": (0, 0, 1)}.get(c.tag, tuple(v[1:4]) if len(v) >= 4 else (0, 0, 1))
            a = np.asarray(axis, float)
            a = a / np.linalg.norm(a)
            th = math.radians(v[0])
            k = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            r = np.eye(3) + math.sin(th) * k + (1 - math.cos(th)) * (k @ k)      # column-vector rotation
            t[0:3, 0:3] = r.T
        else:
            continue
        m = m @ t
    return m


def load_egg(path) -> MeshData:
    with open(path, "r", encoding="utf-8", errors="replace") as fh:
        root = _egg_parse(fh.read())
    y_up = False
    cs = root.child("coordinatesystem")
    if cs is not None and cs.values:
        y_up = cs.values[0].lower().replace("_", "-") in ("y-up", "y-up-right", "yup")

    pools: dict[str, dict[int, tuple]] = {}

    def collect(node):
        for c in node.children:
            if c.tag == "vertexpool":
                pool = {}
                for v in c.children:
                    if v.tag != "vertex":
                        continue
                    xyz = (v.floats() + [0.0, 0.0, 0.0])[:3]
                    nn = v.child("normal")
                    uu = v.child("uv")
                    pool[int(v.name)] = (np.array(xyz), None if nn is None else np.array(nn.floats()[:3]),
                                         None if uu is None else np.array(uu.floats()[:2]))
                pools[c.name] = pool
            else:
                collect(c)

    collect(root)
    out = _Builder()
    state = {"two_sided": False}

    def emit(node, xf):
        for c in node.children:
            if c.tag in ("group", "instance"):
                sub = xf
                tr = c.child("transform")
                if c.tag == "instance" and tr is not None:
                    sub = _egg_transform(tr) @ xf
                emit(c, sub)
            elif c.tag == "polygon":
                ref = c.child("vertexref")
                if ref is None:
                    continue
                pr = ref.child("ref")
                if pr is None or not pr.values or pr.values[0] not in pools:
                    raise ValueError(f"{path}: <VertexRef> names an unknown vertex pool")
                pool = pools[pr.values[0]]
                verts = [pool[int(i)] for i in ref.values]
                if len(verts) < 3:
                    continue
                bf = c.child("bface")
                if bf is not None and bf.values and bf.values[0] not in ("0", "false"):
                    state["two_sided"] = True
                pts = np.array([v[0] for v in verts])
                pts = (np.c_[pts, np.ones(len(pts))] @ xf)[:, :3]
                lin = xf[0:3, 0:3]
                pn = c.child("normal")
                face_n = np.array(pn.floats()[:3]) if pn is not None else None
                geo_n = None
                ids = []
                for (p0, vn, uv), p in zip(verts, pts):
                    n = vn if vn is not None else face_n
                    if n is None:
                        if geo_n is None:
                            geo_n = _face_normal(pts)
                        n = geo_n
                    else:
                        n = n @ np.linalg.inv(lin).T if not np.allclose(lin, np.eye(3)) else n
                    if y_up:
                        p = np.array([p[0], -p[2], p[1]])
                        n = np.array([n[0], -n[2], n[1]])
                    ids.append(out.corner(p, n, uv))
                for k in range(1, len(ids) - 1):
                    out.tris.append((ids[0], ids[k], ids[k + 1]))

    emit(root, np.eye(4))
    return out.mesh(state["two_sided"])


def write_egg(path, mesh: MeshData, polygons=None, comment: str | None = None) -> None:
    """Write ``mesh`` as a Z-up .egg.  ``polygons``: optional list of corner-index lists (n-gons)
    that replaces ``mesh.idx``."""
    faces = [list(map(int, t)) for t in mesh.idx] if polygons is None else [list(map(int, p)) for p in polygons]
    w = ["<CoordinateSystem> { Z-Up }", ""]
    if comment:
        w += ["<Comment> {", f'  "{comment}"', "}"]
    w.append("<VertexPool> vpool {")
    for i, (p, n) in enumerate(zip(mesh.pos, mesh.nrm)):
        w.append(f"  <Vertex> {i} {{")
        w.append("    " + " ".join(f"{float(c):.9g}" for c in p))
        w.append("    <Normal> { " + " ".join(f"{float(c):.9g}" for c in n) + " }")
        if mesh.uv is not None:
            w.append("    <UV> { " + " ".join(f"{float(c):.9g}" for c in mesh.uv[i]) + " }")
        w.append("  }")
    w.append("}")
    w.append("<Group> mesh {")
    for f in faces:
        w.append("  <Polygon> {")
        if mesh.two_sided:
            w.append("    <BFace> { 1 }")
        w.append("    <VertexRef> { " + " ".join(str(i) for i in f) + " <Ref> { vpool } }")
        w.append("  }")
    w.append("}")
    with open(path, "w", encoding="utf-8") as fh:
        fh.write("\n".join(w) + "\n")


# --------------------------------------------------------------------------------------------- glTF
_GLTF_DTYPE = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_GLTF_COUNT = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}


def _gltf_document(path):
    with open(path, "rb") as fh:
        raw = fh.read()
    glb_bin = None
    if raw[:4] == b"glTF":
        _magic, _ver, total = struct.unpack_from("<4sII", raw, 0)
        off, doc = 12, None
        while off + 8 <= min(total, len(raw)):
            clen, ctype = struct.unpack_from("<II", raw, off)
            chunk = raw[off + 8: off + 8 + clen]
            if ctype == 0x4E4F534A:
                doc = json.loads(chunk.decode("utf-8"))
            elif ctype == 0x004E4942 and glb_bin is None:
                glb_bin = chunk
            off += 8 + clen + ((-clen) % 4)
        if doc is None:
            raise ValueError(f"{path}: GLB without a JSON chunk")
    else:
        doc = json.loads(raw.decode("utf-8"))
    base = os.path.dirname(os.path.abspath(path))
    buffers = []
    for b in doc.get("buffers", []):
        uri = b.get("uri")
        if uri is None:
            if glb_bin is None:
                raise ValueError(f"{path}: buffer without uri outside a GLB")
            buffers.append(glb_bin)
        elif uri.startswith("data:"):
            buffers.append(base64.b64decode(uri.split(",", 1)[1]))
        else:
            from urllib.parse import unquote
            with open(os.path.join(base, unquote(uri)), "rb") as fh:
                buffers.append(fh.read())
    return doc, buffers


def _gltf_accessor(doc, buffers, index) -> np.ndarray:
    acc = doc["accessors"][index]
    dt = np.dtype(_GLTF_DTYPE[acc["componentType"]])
    ncomp = _GLTF_COUNT[acc["type"]]
    count = acc["count"]
    if "bufferView" not in acc:
        return np.zeros((count, ncomp), dtype=dt)
    bv = doc["bufferViews"][acc["bufferView"]]
    start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
    item = dt.itemsize * ncomp
    stride = bv.get("byteStride") or item
    data = buffers[bv["buffer"]]
    if stride == item:
        arr = np.frombuffer(data, dtype=dt, count=count * ncomp, offset=start).reshape(count, ncomp)
    else:
        rows = np.frombuffer(data, dtype=np.uint8, count=(count - 1) * stride + item, offset=start) if count else np.zeros(0, np.uint8)
        arr = np.lib.stride_tricks.as_strided(rows, shape=(count, item), strides=(stride, 1)).copy().view(dt).reshape(count, ncomp)
    if acc.get("normalized") and dt.kind in "iu":
        info = np.iinfo(dt)
        arr = np.maximum(arr.astype(np.float32) / float(info.max), -1.0)
    return arr


def _gltf_node_matrix(node) -> np.ndarray:
    if "matrix" in node:
        return np.array(node["matrix"], dtype=np.float64).reshape(4, 4).T          # stored column-major
    m = np.eye(4)
    if "scale" in node:
        m = np.diag(list(node["scale"]) + [1.0]) @ m
    if "rotation" in node:
        x, y, z, w = node["rotation"]
        r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        r4 = np.eye(4)
        r4[0:3, 0:3] = r
        m = r4 @ m
    if "translation" in node:
        t = np.eye(4)
        t[0:3, 3] = node["translation"]
        m = t @ m
    return m


def load_gltf(path) -> MeshData:
    doc, buffers = _gltf_document(path)
    out = _Builder()
    two_sided = False
    scenes = doc.get("scenes") or [{"nodes": list(range(len(doc.get("nodes", []))))}]
    roots = scenes[doc.get("scene", 0)].get("nodes", [])

    def prim_triangles(prim, nverts):
        idx = (_gltf_accessor(doc, buffers, prim["indices"]).reshape(-1).astype(np.int64)
               if "indices" in prim else np.arange(nverts, dtype=np.int64))
        mode = prim.get("mode", 4)
        if mode == 4:
            return idx[: (len(idx) // 3) * 3].reshape(-1, 3)
        if mode == 5:        # strip: winding alternates
            t = [(idx[i], idx[i + 1], idx[i + 2]) if i % 2 == 0 else (idx[i + 1], idx[i], idx[i + 2]) for i in range(len(idx) - 2)]
            return np.array(t, dtype=np.int64).reshape(-1, 3)
        if mode == 6:        # fan
            return np.array([(idx[0], idx[i], idx[i + 1]) for i in range(1, len(idx) - 1)], dtype=np.int64).reshape(-1, 3)
        return np.zeros((0, 3), dtype=np.int64)          # points / lines: nothing to rasterise

    def visit(ni, parent):
        nonlocal two_sided
        node = doc["nodes"][ni]
        world = parent @ _gltf_node_matrix(node)
        if "mesh" in node:
            lin = world[0:3, 0:3]
            nmat = np.linalg.inv(lin).T if abs(np.linalg.det(lin)) > 1e-30 else lin
            flip = np.linalg.det(lin) < 0
            for prim in doc["meshes"][node["mesh"]].get("primitives", []):
                attrs = prim.get("attributes", {})
                if "POSITION" not in attrs:
                    continue
                pos = _gltf_accessor(doc, buffers, attrs["POSITION"]).astype(np.float64)
                nrm = _gltf_accessor(doc, buffers, attrs["NORMAL"]).astype(np.float64) if "NORMAL" in attrs else None
                uv = _gltf_accessor(doc, buffers, attrs["TEXCOORD_0"]).astype(np.float64) if "TEXCOORD_0" in attrs else None
                mat = prim.get("material")
                if mat is not None and doc.get("materials", [])[mat].get("doubleSided"):
                    two_sided = True
                wp = pos @ lin.T + world[0:3, 3]
                wn = None
                if nrm is not None:
                    wn = nrm @ nmat.T
                    ln = np.linalg.norm(wn, axis=1, keepdims=True)
                    wn = wn / np.where(ln > 0, ln, 1.0)
                # Y-up -> Z-up
                wp = np.stack([wp[:, 0], -wp[:, 2], wp[:, 1]], axis=1)
                if wn is not None:
                    wn = np.stack([wn[:, 0], -wn[:, 2], wn[:, 1]], axis=1)
                for a, b, c in prim_triangles(prim, len(pos)):
                    if flip:
                        b, c = c, b
                    fn = None if wn is not None else _face_normal(wp[[a, b, c]])
                    ids = [out.corner(wp[i], wn[i] if wn is not None else fn, None if uv is None else (uv[i, 0], 1.0 - uv[i, 1]))
                           for i in (a, b, c)]
                    out.tris.append(tuple(ids))
        for ch in node.get("children", []):
            visit(ch, world)

    for r in roots:
        visit(r, np.eye(4))
    return out.mesh(two_sided)


def write_gltf(path, mesh: MeshData, node_matrices=(), y_up_source: bool = True, asset_extras=None) -> None:
    """Write ``mesh`` (Z-up, as :func:`load_gltf` returns it) as ``path`` + a sibling ``.bin``.

    ``node_matrices``: 4x4 matrices (column-vector convention) of wrapper nodes above the mesh node;
    the vertex data is stored pre-multiplied by their inverse so that loading reproduces ``mesh``.
    """
    pos = mesh.pos.astype(np.float64)
    nrm = mesh.nrm.astype(np.float64)
    if y_up_source:          # inverse of (x, y, z) -> (x, -z, y)
        pos = np.stack([pos[:, 0], pos[:, 2], -pos[:, 1]], axis=1)
        nrm = np.stack([nrm[:, 0], nrm[:, 2], -nrm[:, 1]], axis=1)
    world = np.eye(4)
    for m in node_matrices:
        world = world @ np.asarray(m, dtype=np.float64)
    inv = np.linalg.inv(world)
    pos = pos @ inv[0:3, 0:3].T + inv[0:3, 3]
    nrm = nrm @ world[0:3, 0:3]              # (inv^-T)^T = world^T ... normals use inverse transpose of inv
    pos32, nrm32 = pos.astype(np.float32), nrm.astype(np.float32)
    idx = mesh.idx.astype(np.uint32).reshape(-1)
    blob_idx = idx.tobytes()
    blob_uv = b"" if mesh.uv is None else np.stack([mesh.uv[:, 0], 1.0 - mesh.uv[:, 1]], axis=1).astype(np.float32).tobytes()
    blob_pn = pos32.tobytes() + nrm32.tobytes()
    views = [{"buffer": 0, "byteLength": len(blob_idx), "target": 34963}]
    off = len(blob_idx)
    uv_view = None
    if blob_uv:
        uv_view = len(views)
        views.append({"buffer": 0, "byteOffset": off, "byteLength": len(blob_uv), "byteStride": 8, "target": 34962})
        off += len(blob_uv)
    pn_view = len(views)
    views.append({"buffer": 0, "byteOffset": off, "byteLength": len(blob_pn), "byteStride": 12, "target": 34962})
    nv = len(pos32)
    accessors = [
        {"bufferView": pn_view, "componentType": 5126, "count": nv, "type": "VEC3",
         "min": [float(v) for v in pos32.min(0)], "max": [float(v) for v in pos32.max(0)]},
        {"bufferView": pn_view, "byteOffset": nv * 12, "componentType": 5126, "count": nv, "type": "VEC3"},
        {"bufferView": 0, "componentType": 5125, "count": int(idx.size), "type": "SCALAR"},
    ]
    attributes = {"POSITION": 0, "NORMAL": 1}
    if uv_view is not None:
        attributes["TEXCOORD_0"] = len(accessors)
        accessors.append({"bufferView": uv_view, "componentType": 5126, "count": nv, "type": "VEC2"})
    nodes = []
    for i, m in enumerate(node_matrices):
        nodes.append({"name": f"wrapper_{i}", "children": [i + 1],
                      "matrix": [float(v) for v in np.asarray(m, dtype=np.float64).T.reshape(-1)]})
    nodes.append({"name": "mesh", "mesh": 0})
    bin_name = os.path.splitext(os.path.basename(path))[0] + ".bin"
    doc = {
        "asset": {"version": "2.0", "generator": "pybatchrender_b200.mesh_io", **({"extras": asset_extras} if asset_extras else {})},
        "scene": 0,
        "scenes": [{"nodes": [0]}],
        "nodes": nodes,
        "meshes": [{"primitives": [{"attributes": attributes, "indices": 2, "material": 0, "mode": 4}]}],
        "materials": [{"doubleSided": bool(mesh.two_sided), "pbrMetallicRoughness": {"metallicFactor": 0.0, "roughnessFactor": 0.6}}],
        "accessors": accessors,
        "bufferViews": views,
        "buffers": [{"byteLength": len(blob_idx) + len(blob_uv) + len(blob_pn), "uri": bin_name}],
    }
    with open(os.path.join(os.path.dirname(os.path.abspath(path)), bin_name), "wb") as fh:
        fh.write(blob_idx + blob_uv + blob_pn)
    with open(path, "w", encoding="utf-8") as fh:
        json.dump(doc, fh, indent=2)
        fh.write("\n")
