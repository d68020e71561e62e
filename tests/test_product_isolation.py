"""The oracle is test infrastructure: nothing under pybatchrender_b200/ may import, load or execute
anything under oracle/ (a product path routed through the oracle would void every parity claim)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_oracle_reference_in_product_sources():
    bad = []
    for dp, _dn, fns in os.walk(os.path.join(ROOT, "pybatchrender_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, fn), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|libpbr_oracle|orc_render", src, flags=re.M):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_importing_the_product_does_not_load_the_oracle():
    code = ("import sys; sys.path.insert(0, %r); import pybatchrender_b200, pybatchrender_b200.envs; "
            "from pybatchrender_b200.envs.cartpole import CartPoleRenderer; "
            "CartPoleRenderer(dict(num_scenes=2, device='cpu')); "
            "assert 'oracle' not in sys.modules; "
            "assert 'libpbr_oracle' not in open('/proc/self/maps').read(); print('ok')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
