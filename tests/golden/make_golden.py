#!/usr/bin/env python
"""Regenerate the golden fixtures in this directory from the reference tree.

Run in the authoring container only (needs /root/reference, PIL):

    python tests/golden/make_golden.py

Sources (all inside the read-only reference checkout):

* ``examples/notebooks/cartpole_benchmark.ipynb`` cell 14 output ``image/png``
  (raw file line 473): a 768x768 RGB grid written by
  ``env.save_batch_examples(pixels=td["pixels"], num=16, scale=3)``
  (reference ``pybatchrender/env.py:97-244``).  Grid logic: tile n sits at row
  n//4, col n%4, each tile nearest-neighbour upscaled x3, so ``[::3, ::3]``
  recovers the 16 rendered 64x64 tiles bit-exactly.  These are scenes 0..15 of
  ``pbr.envs.make("CartPole-v0", num_scenes=4098, tile_resolution=(64, 64))``
  right after ``reset()`` (ipynb cell 10/12), rendered by Panda3D on an NVIDIA L4.
* cell 12 text output (raw lines 422-424): observation[0:3] of the same reset.
* cell 20 output ``image/png`` (raw line 670): the same grid after the benchmark
  loop (states not printed; used for the rail mask and the colour palette only).

Outputs: ``cartpole_nb.npz`` with ``initial`` [16,64,64,3] u8, ``final``
[16,64,64,3] u8, ``obs3`` [3,4] f32 (as printed, 4 decimals).
"""
import base64
import io
import json
import os

import numpy as np
from PIL import Image

REF_NB = "/root/reference/examples/notebooks/cartpole_benchmark.ipynb"
HERE = os.path.dirname(os.path.abspath(__file__))


def _grid_png(cell):
    for o in cell["outputs"]:
        if "data" in o and "image/png" in o["data"]:
            im = Image.open(io.BytesIO(base64.b64decode(o["data"]["image/png"]))).convert("RGB")
            a = np.array(im)
            assert a.shape == (768, 768, 3), a.shape
            s = a[::3, ::3]
            for i in range(3):
                for j in range(3):
                    assert (a[i::3, j::3] == s).all(), "grid is not an exact x3 nearest upscale"
            tiles = s.reshape(4, 64, 4, 64, 3).transpose(0, 2, 1, 3, 4).reshape(16, 64, 64, 3)
            return np.ascontiguousarray(tiles)
    raise RuntimeError("no png in cell")


def main():
    nb = json.load(open(REF_NB))
    initial = _grid_png(nb["cells"][14])
    final = _grid_png(nb["cells"][20])
    obs3 = np.array(
        [[-0.4038, 0.9445, -0.5032, 0.1505],
         [0.0667, 0.5819, 0.2862, -0.1914],
         [-1.9003, -0.0619, 0.3823, -0.1472]], dtype=np.float32)
    txt = "".join("".join(o.get("text", [])) for o in nb["cells"][12]["outputs"])
    assert "-0.4038,  0.9445, -0.5032,  0.1505" in txt and "-1.9003, -0.0619,  0.3823, -0.1472" in txt
    out = os.path.join(HERE, "cartpole_nb.npz")
    np.savez_compressed(out, initial=initial, final=final, obs3=obs3)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
