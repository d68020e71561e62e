"""PBRConfig arithmetic (reference config.py:61-145) -- property tests."""
import math

import pytest
from hypothesis import given, settings, strategies as st

from pybatchrender_b200 import PBRConfig
from pybatchrender_b200.config import grid_for


@given(st.integers(min_value=1, max_value=200000))
@settings(max_examples=200, deadline=None)
def test_grid_holds_all_scenes_and_is_near_square(n):
    cols, rows = grid_for(n)
    assert cols * rows >= n and cols * (rows - 1) < n
    assert cols == math.ceil(math.sqrt(n))


@given(st.integers(1, 5000), st.integers(1, 300), st.integers(1, 300))
@settings(max_examples=100, deadline=None)
def test_window_is_tiles_times_tile(n, w, h):
    c = PBRConfig(num_scenes=n, tile_resolution=(w, h), device="cpu")
    assert c.window_resolution == (c.tiles[0] * w, c.tiles[1] * h)
    assert c.batch_inner_dim == c.tiles[0] * c.tiles[1] >= c.num_scenes == n


def test_defaults_and_fill_in_rules():
    c = PBRConfig(device="cpu")
    assert (c.tiles, c.tile_resolution, c.window_resolution, c.num_scenes) == ((1, 1), (64, 64), (64, 64), 1)
    c = PBRConfig(num_scenes=4096, device="cpu")
    assert c.tiles == (64, 64) and c.window_resolution == (4096, 4096)
    c = PBRConfig(num_scenes=4098, device="cpu")
    assert c.tiles == (65, 64)
    c = PBRConfig(tiles=(3, 2), window_resolution=(96, 40), device="cpu")
    assert c.tile_resolution == (32, 20) and c.num_scenes == 6
    c = PBRConfig(tile_resolution=(10, 20), window_resolution=(40, 40), device="cpu")
    assert c.tiles == (4, 2)
    c = PBRConfig(window_resolution=(100, 50), device="cpu")
    assert c.tiles == (1, 1) and c.tile_resolution == (100, 50)


def test_errors():
    with pytest.raises(ValueError):
        PBRConfig(tiles=(2, 2), tile_resolution=(8, 8), window_resolution=(17, 16), device="cpu")
    with pytest.raises(ValueError):
        PBRConfig(num_scenes=5, tiles=(2, 2), device="cpu")
    with pytest.raises(ValueError):
        PBRConfig(num_scenes=4, batch_inner_dim=5, device="cpu")
    with pytest.raises(ValueError):
        PBRConfig(device="tpu")


def test_from_config_variants_and_prc():
    base = PBRConfig(num_scenes=9, device="cpu")
    c = PBRConfig.from_config(base, num_channels=4)
    assert c.num_scenes == 9 and c.num_channels == 4 and c.tiles == (3, 3)
    c = PBRConfig.from_config({"num_scenes": 2, "device": "cpu"})
    assert c.num_scenes == 2
    assert PBRConfig.from_config(None, device="cpu").num_scenes == 1
    with pytest.raises(TypeError):       # quirk Q7: worker_index is not a PBRConfig field
        PBRConfig.from_config(base, worker_index=1)
    prc = base.build_prc()
    assert "window-type offscreen" in prc and "win-size 192 192" in prc
    assert "num_scenes: 9" in repr(base)
