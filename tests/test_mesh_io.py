"""Model readers: .egg and glTF subsets against procedural geometry, the
bundled generated assets, and -- in the authoring container -- the reference's own model files."""
import base64
import json
import os
import struct

import numpy as np
import pytest

from pybatchrender_b200 import mesh_io, meshes

REF_MODELS = "/root/reference/pybatchrender/models"


def _tri_set(m, decimals=4):
    """Order-independent description of the triangles: sorted rows of rounded corner positions."""
    t = m.pos[m.idx.astype(np.int64)]                       # [T,3,3]
    rows = []
    for tri in np.round(t.astype(np.float64), decimals) + 0.0:
        k = min(range(3), key=lambda i: tuple(tri[i]))       # rotate: smallest corner first (keeps winding)
        rows.append(tuple(np.roll(tri, -k, axis=0).reshape(-1)))
    return sorted(rows)


def _outward(m):
    t = m.pos[m.idx.astype(np.int64)].astype(np.float64)
    g = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    n = m.nrm[m.idx[:, 0].astype(np.int64)].astype(np.float64)
    return np.einsum("ij,ij->i", g, n)


def test_egg_round_trip(tmp_path):
    for name, m in (("box", meshes.box()), ("sphere", meshes.uv_sphere(1.5, 8, 5)), ("cone", meshes.cone(12))):
        p = tmp_path / f"{name}.egg"
        mesh_io.write_egg(p, m)
        got = mesh_io.load_egg(p)
        assert got.n_tris == m.n_tris
        assert _tri_set(got) == _tri_set(m)
        # normals survive per corner
        a = np.round(got.nrm[got.idx.astype(np.int64)].reshape(-1, 3), 5)
        assert np.isfinite(a).all() and (_outward(got) > 0).all()


def test_gltf_round_trip_with_wrapper_nodes(tmp_path):
    rot = np.array([[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=float)
    shift = np.eye(4); shift[0:3, 3] = (1.0, -2.0, 0.5)
    for name, m in (("box", meshes.box()), ("cyl", meshes.cylinder(2.0, 3.0, 8, 2))):
        p = tmp_path / name / "scene.gltf"
        os.makedirs(p.parent)
        mesh_io.write_gltf(p, m, node_matrices=(rot, shift))
        got = mesh_io.load_gltf(p)
        assert got.two_sided == m.two_sided
        assert _tri_set(got, 3) == _tri_set(m, 3)
        gn = got.nrm[got.idx.astype(np.int64)].reshape(-1, 3)
        assert np.allclose(np.linalg.norm(gn, axis=1), 1.0, atol=1e-5)
        assert (_outward(got) > 0).all()


def test_bundled_models_resolve_like_the_reference_paths():
    c = meshes.load_mesh("models/cone.egg")
    assert c.n_tris == 62 and not c.two_sided                   # 30 cap-fan + 32 side triangles
    assert np.allclose(c.pos.min(0), (-1, -1, -1)) and np.allclose(c.pos.max(0), (1, 1, 1))
    assert (_outward(c) > 0).all()
    cap = np.isclose(c.nrm[:, 1], -1.0)
    assert cap.sum() == 32 and np.allclose(c.pos[cap, 1], -1.0)
    y = meshes.load_mesh("models/cylinder/scene.gltf")
    assert y.n_tris == 320 and y.pos.shape[0] == 226 and y.two_sided
    assert np.allclose(y.pos.min(0), (-50, -50, -100), atol=1e-3) and np.allclose(y.pos.max(0), (50, 50, 100), atol=1e-3)
    assert (_outward(y) > 0).all()
    assert _tri_set(y, 2) == _tri_set(meshes.cylinder(), 2)
    with pytest.raises(FileNotFoundError):
        meshes.load_mesh("models/does_not_exist.egg")


def test_egg_features(tmp_path):
    p = tmp_path / "f.egg"
    p.write_text("""
<CoordinateSystem> { Y-Up }
// a comment
<VertexPool> vp {
  <Vertex> 1 { 0 0 0 <UV> { 0 0 } }
  <Vertex> 2 { 1 0 0 <UV> { 1 0 } }
  <Vertex> 3 { 1 1 0 <UV> { 1 1 } }
  <Vertex> 4 { 0 1 0 <UV> { 0 1 } }
}
<Group> g {
  <Polygon> { <BFace> { 1 } <VertexRef> { 1 2 3 4 <Ref> { vp } } }
  <Instance> moved {
    <Transform> { <Scale> { 2 } <Translate> { 0 0 5 } }
    <Polygon> { <Normal> { 0 0 1 } <VertexRef> { 1 2 3 <Ref> { vp } } }
  }
}
""")
    m = mesh_io.load_egg(p)
    assert m.two_sided and m.n_tris == 3 and m.uv is not None
    # quad in the Y-up xy-plane, normal +z (Y-up)  ->  Z-up: (x, -z, y), normal (0, -1, 0)
    quad = m.pos[m.idx[:2].astype(np.int64)].reshape(-1, 3)
    assert np.allclose(quad[:, 1], 0.0) and np.allclose(m.nrm[m.idx[0, 0]], (0, -1, 0))
    # instanced triangle: scaled by 2 then moved 5 along Y-up z  ->  Z-up y = -5
    tri = m.pos[m.idx[2].astype(np.int64)]
    assert np.allclose(tri[:, 1], -5.0) and np.allclose(tri[:, 0].max(), 2.0) and np.allclose(tri[:, 2].max(), 2.0)
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.egg"
        bad.write_text("<VertexPool> vp { <Vertex> 0 { 0 0 0 }")
        mesh_io.load_egg(bad)


def _tiny_gltf(tmp_path, *, glb=False, data_uri=False, u16=True, normals=False, mode=4, trs=True):
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3] if mode == 4 else [0, 1, 3, 2], np.uint16 if u16 else np.uint32)
    blob = idx.tobytes()
    blob += b"\0" * ((-len(blob)) % 4)
    pos_off = len(blob)
    blob += pos.tobytes()
    doc = {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0]}],
           "nodes": [{"mesh": 0, **({"translation": [0, 2, 0], "rotation": [0, 0, 0.70710678, 0.70710678], "scale": [2, 2, 2]} if trs else {})}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1, "mode": mode}]}],
           "accessors": [{"bufferView": 1, "componentType": 5126, "count": 4, "type": "VEC3"},
                         {"bufferView": 0, "componentType": 5123 if u16 else 5125, "count": int(idx.size), "type": "SCALAR"}],
           "bufferViews": [{"buffer": 0, "byteLength": idx.nbytes}, {"buffer": 0, "byteOffset": pos_off, "byteLength": pos.nbytes}],
           "buffers": [{"byteLength": len(blob)}]}
    if normals:
        nb = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1)).tobytes()
        doc["bufferViews"].append({"buffer": 0, "byteOffset": len(blob), "byteLength": len(nb)})
        doc["accessors"].append({"bufferView": 2, "componentType": 5126, "count": 4, "type": "VEC3"})
        doc["meshes"][0]["primitives"][0]["attributes"]["NORMAL"] = 2
        blob += nb
        doc["buffers"][0]["byteLength"] = len(blob)
    if glb:
        js = json.dumps(doc).encode()
        js += b" " * ((-len(js)) % 4)
        body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
        p = tmp_path / "t.glb"
        p.write_bytes(struct.pack("<4sII", b"glTF", 2, 12 + len(body)) + body)
        return p
    if data_uri:
        doc["buffers"][0]["uri"] = "data:application/octet-stream;base64," + base64.b64encode(blob).decode()
    else:
        (tmp_path / "t.bin").write_bytes(blob)
        doc["buffers"][0]["uri"] = "t.bin"
    p = tmp_path / "t.gltf"
    p.write_text(json.dumps(doc))
    return p


@pytest.mark.parametrize("kw", [dict(), dict(glb=True), dict(data_uri=True), dict(u16=False, normals=True),
                                dict(mode=5), dict(trs=False, normals=True)])
def test_gltf_features(tmp_path, kw):
    m = mesh_io.load_gltf(_tiny_gltf(tmp_path, **kw))
    assert m.n_tris == 2 and not m.two_sided
    if kw.get("trs", True):
        # glTF: scale 2, rotate +90 deg about z, translate y+2: (x,y,0) -> (-2y, 2x+2, 0); Z-up: (x,-z,y)
        want = {(0.0, 0.0, 2.0), (0.0, 0.0, 4.0), (-2.0, 0.0, 4.0), (-2.0, 0.0, 2.0)}
    else:
        want = {(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 0.0, 1.0), (0.0, 0.0, 1.0)}
    got = {tuple(float(v) for v in np.round(p, 4) + 0.0) for p in m.pos}
    assert got == want
    # face normal: glTF +z  ->  Z-up -y; winding stays counter-clockwise around it
    assert np.allclose(m.nrm, (0, -1, 0), atol=1e-6) and (_outward(m) > 0).all()


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference checkout not present")
def test_reference_assets_load_and_match_the_procedural_shapes():
    c = mesh_io.load_file(os.path.join(REF_MODELS, "cone.egg"))
    mine = meshes.cone(32)
    assert c.n_tris == 62 and (_outward(c) > 0).all()
    assert _tri_set(c, 4) == _tri_set(mine, 4)
    # side normals agree away from the apex (the asset's apex normals are obj2egg smoothing-group averages)
    ref_n = {tuple(np.round(p, 3)): n for p, n in zip(c.pos, c.nrm) if p[1] < 0 and n[1] > 0}
    for p, n in zip(mine.pos, mine.nrm):
        if p[1] < 0 and n[1] > 0:
            assert np.allclose(ref_n[tuple(np.round(p, 3))], n, atol=2e-5)
    y = mesh_io.load_file(os.path.join(REF_MODELS, "cylinder", "scene.gltf"))
    assert y.n_tris == 320 and y.pos.shape[0] == 226 and y.two_sided
    assert (_outward(y) > 0).all()
    assert _tri_set(y, 2) == _tri_set(meshes.cylinder(), 2)
    cyl = meshes.cylinder()
    key = lambda p, n: tuple(np.round(np.r_[p, 10.0 * n], 1) + 0.0)
    assert sorted(key(p, n) for p, n in zip(y.pos, y.nrm)) == sorted(key(p, n) for p, n in zip(cyl.pos, cyl.nrm))
