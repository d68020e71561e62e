"""The C ABI: every function declared in include/pbr_b200.h is exported by libpbr_b200.so with C
linkage, the library loads without a GPU, and the Python binding knows every symbol."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pbr_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"#ifdef PBR_W_TIMING.*?#endif", "", src, flags=re.S)      # timing builds only (profiles/)
    return sorted(set(re.findall(r"\b(pbr_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    from pybatchrender_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _native.LIB_PATH


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ["pbr_render", "pbr_mesh_create", "pbr_mesh_destroy", "pbr_compose_transforms",
                 "pbr_pack_transforms", "pbr_last_error", "pbr_version"]:
        assert must in names


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (pbr_[a-z_0-9]+)\b", out))
    assert set(declared_functions()) <= exported, set(declared_functions()) - exported


def test_library_loads_and_reports_version_without_gpu(libpath):
    lib = ctypes.CDLL(libpath)
    lib.pbr_version.restype = ctypes.c_int
    assert lib.pbr_version() >= 100
    lib.pbr_last_error.restype = ctypes.c_char_p
    assert lib.pbr_last_error() is not None


def test_python_binding_covers_the_header(libpath):
    from pybatchrender_b200 import _native
    assert set(declared_functions()) == set(_native.exported_symbols())
    _native.load()


def test_argument_validation_without_gpu(libpath):
    """Validation that happens before any CUDA call can be exercised on the CPU box."""
    from pybatchrender_b200 import _native
    lib = _native.load()
    assert lib.pbr_render(None, None) == -1
    assert b"NULL frame" in lib.pbr_last_error()
    f = _native._FrameDesc()
    f.tile_w, f.tile_h, f.channels = 0, 64, 3
    assert lib.pbr_render(ctypes.byref(f), None) == -1
    f.tile_w, f.channels = 64, 5
    assert lib.pbr_render(ctypes.byref(f), None) == -1
    handle = ctypes.c_void_p()
    assert lib.pbr_mesh_create(None, None, None, 0, None, 0, 0, 0, ctypes.byref(handle)) == -1


def test_kernels_are_sm100a_only(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
