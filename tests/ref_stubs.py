"""Run the reference's own host-side classes without Panda3D.

The reference's ``PBRConfig``, ``PBRCam``, ``PBRNode``, ``PBRLight`` and ``PBRRenderer._rearrange_img``
are pure torch; only their imports and buffer-texture plumbing touch Panda3D.  This module installs
fake ``panda3d`` / ``direct`` modules (a ``Texture`` that records ``set_ram_image`` bytes, a
recording ``NodePath``), puts ``/root/reference`` on ``sys.path`` and hands back the reference
classes plus a fake ``ShowBase``.  It exists only in the authoring container (the GPU box has no
/root/reference); tests that use it are skipped elsewhere and the vectors it produces are committed
as ``tests/golden/host_math.npz`` (see ``tests/golden/make_host_golden.py``).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "pybatchrender"))


class _Texture:
    T_float = 1
    F_rgba32 = 2
    CMOff = 0

    def __init__(self, name=""):
        self.name = name
        self.ram = b""
        self.texels = 0

    def setupBufferTexture(self, texels, *_a):
        self.texels = int(texels)

    def set_keep_ram_image(self, *_a):
        pass

    def set_ram_image(self, data):
        self.ram = bytes(data)


class _NodePath:
    def __init__(self, name="np"):
        self.name = name
        self.inputs = {}
        self.calls = []

    def __getattr__(self, item):
        def _rec(*a, **k):
            self.calls.append((item, a, k))
            return None
        return _rec

    def setShaderInput(self, name, value):
        self.inputs[name] = value

    def getTightBounds(self, *_a):
        return None

    def node(self):
        return self

    def attachNewNode(self, name):
        return _NodePath(name)

    def getParent(self):
        return _NodePath("parent")


class _Loader:
    def loadModel(self, path):
        return _NodePath(str(path))


class _Win:
    def __init__(self, x, y):
        self.x, self.y = x, y

    def getXSize(self):
        return self.x

    def getYSize(self):
        return self.y


class FakeBase:
    def __init__(self, win_x, win_y):
        self.win = _Win(win_x, win_y)
        self.loader = _Loader()
        self.render = _NodePath("render")
        self._pbr_nodes = []
        self._pbr_cam = None
        self._pbr_light = None


def install():
    """-> dict(PBRConfig, PBRCam, PBRNode, PBRLight, PBRShaderContext, rearrange) from the reference."""
    if not available():
        raise RuntimeError("reference checkout not present")
    if "panda3d" not in sys.modules or not getattr(sys.modules["panda3d"], "_pbr_stub", False):
        panda = types.ModuleType("panda3d")
        panda._pbr_stub = True
        core = types.ModuleType("panda3d.core")
        core.Texture = _Texture
        core.NodePath = _NodePath
        core.OmniBoundingVolume = lambda *a, **k: None
        core.Shader = types.SimpleNamespace(make=lambda *a, **k: None, SL_GLSL=0)
        core.GeomEnums = types.SimpleNamespace(UH_dynamic=0)
        core.loadPrcFileData = lambda *a, **k: None
        core.MouseButton = types.SimpleNamespace(one=lambda: 0)
        core.TextNode = types.SimpleNamespace(ALeft=0)
        core.LPoint3 = lambda *a: a
        core.Vec3 = lambda *a: a
        panda.core = core
        direct = types.ModuleType("direct")
        showbase = types.ModuleType("direct.showbase")
        sb = types.ModuleType("direct.showbase.ShowBase")
        sb.ShowBase = type("ShowBase", (), {"__init__": lambda self, *a, **k: None})
        sbg = types.ModuleType("direct.showbase.ShowBaseGlobal")
        sbg.globalClock = types.SimpleNamespace(getRealTime=lambda: 0.0, getAverageFrameRate=lambda: 0.0)
        task = types.ModuleType("direct.task")
        task.Task = types.SimpleNamespace(cont=0, done=1)
        gui = types.ModuleType("direct.gui")
        ost = types.ModuleType("direct.gui.OnscreenText")
        ost.OnscreenText = lambda *a, **k: None
        for name, mod in {"panda3d": panda, "panda3d.core": core, "direct": direct, "direct.showbase": showbase,
                          "direct.showbase.ShowBase": sb, "direct.showbase.ShowBaseGlobal": sbg,
                          "direct.task": task, "direct.gui": gui, "direct.gui.OnscreenText": ost}.items():
            sys.modules[name] = mod
    # load the reference package under a private name: this repo ships its own ``pybatchrender``
    # alias package, which must not shadow (or be shadowed by) the reference checkout
    import importlib
    import importlib.util
    name = "_reference_pybatchrender"
    if name not in sys.modules:
        pkg_dir = os.path.join(REF_ROOT, "pybatchrender")
        spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                      submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    cfg_mod = importlib.import_module(name + ".config")
    sc = importlib.import_module(name + ".renderer.shader_context")
    node = importlib.import_module(name + ".renderer.node")
    cam = importlib.import_module(name + ".renderer.camera")
    light = importlib.import_module(name + ".renderer.light")
    rend = importlib.import_module(name + ".renderer.renderer")
    return dict(PBRConfig=cfg_mod.PBRConfig, PBRShaderContext=sc.PBRShaderContext, PBRNode=node.PBRNode,
                PBRCam=cam.PBRCam, PBRLight=light.PBRLight, rearrange=rend.PBRRenderer._rearrange_img)
