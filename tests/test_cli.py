"""examples/scripts/cartpole_benchmark.py keeps the reference script's command line."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("cartpole_benchmark_cli", os.path.join(ROOT, "examples", "scripts", "cartpole_benchmark.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_flags_parse_and_state_only_run():
    cli = _load()
    a = cli.parse_args(["--num-scenes", "32", "--tile-resolution", "48", "40", "--steps", "3", "--parallel", "--num-workers", "2",
                        "--save-every", "10", "--save-num", "4", "--save-dir", "/tmp/x", "--no-render", "--window",
                        "--gif", "--gif-steps", "5", "--gif-interval", "1", "--gif-scale", "2", "--gif-duration", "40"])
    assert (a.num_scenes, tuple(a.tile_resolution), a.steps, a.parallel, a.offscreen, a.gif) == (32, (48, 40), 3, True, False, True)
    out = cli.run(cli.parse_args(["--no-render", "--device", "cpu", "--num-scenes", "8", "--steps", "4"]))
    assert out["scenes"] == 8 and out["fps"] > 0


@pytest.mark.gpu
def test_render_run_and_gif(tmp_path):
    cli = _load()
    out = cli.run(cli.parse_args(["--num-scenes", "64", "--steps", "5", "--save-every", "2", "--save-dir", str(tmp_path)]))
    assert out["scenes"] == 64 and any(f.endswith(".png") for f in os.listdir(tmp_path))
    g = cli.run(cli.parse_args(["--num-scenes", "16", "--gif", "--gif-steps", "6", "--save-dir", str(tmp_path)]))
    assert os.path.getsize(g["gif"]) > 1000


def test_bench_module_constants_exist():
    """bench.py's GPU arm cannot run here; at least every module-level constant its functions read must
    be defined (a constant once disappeared in an edit and only the GPU box noticed)."""
    import ast
    import builtins
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    local = {"K", "W", "N", "H", "I", "B"}       # one-letter-style locals of bench.py's functions
    missing = {n.id for n in ast.walk(tree)
               if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id.isupper() and len(n.id) > 1
               and not hasattr(bench, n.id) and not hasattr(builtins, n.id) and n.id not in local}
    assert not missing, missing
    assert bench.STATE_RING >= 1 and sorted(bench.CONFIGS) == [2, 3, 4, 5]
    assert bench.CONFIGS[2]["algo"] == 12512 and bench.CONFIGS[4]["algo"] == 21392
    assert bench.CONFIGS[3]["algo"] == 69696 and bench.CONFIGS[5]["algo"] == 201792
