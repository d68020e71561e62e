"""Generate the bundled model files under pybatchrender_b200/models/ from procedural geometry.

The reference ships ``models/cone.egg`` and ``models/cylinder/scene.gltf`` (``envs/steering/config.py:58-59``).
Those files are not copied: equivalent shapes are built by :func:`meshes.cone` / :func:`meshes.cylinder`
and written in the same two formats (the cone with its cap as one n-gon, the cylinder below two
wrapper nodes whose rotations cancel, as exporters often leave them), so that ``add_node("models/cone.egg")``
and ``add_node("models/cylinder/scene.gltf")`` work and exercise the real readers.

    python tools/make_models.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybatchrender_b200 import mesh_io, meshes  # noqa: E402


def main():
    root = meshes.MODELS_DIR
    os.makedirs(os.path.join(root, "cylinder"), exist_ok=True)
    seg = 32
    c = meshes.cone(seg)
    polys = [list(range(seg))] + [list(map(int, t)) for t in c.idx[seg - 2:]]
    mesh_io.write_egg(os.path.join(root, "cone.egg"), c, polygons=polys,
                      comment="procedural cone, 32 segments (tools/make_models.py)")
    rx_m90 = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
    rx_p90 = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
    mesh_io.write_gltf(os.path.join(root, "cylinder", "scene.gltf"), meshes.cylinder(), node_matrices=(rx_m90, rx_p90),
                       asset_extras={"title": "procedural cylinder (tools/make_models.py)"})
    for rel in ("cone.egg", "cylinder/scene.gltf"):
        m = mesh_io.load_file(os.path.join(root, rel))
        print(rel, m.pos.shape, m.idx.shape, m.two_sided, m.pos.min(0), m.pos.max(0))


if __name__ == "__main__":
    main()
