"""``PBRNode`` -- one instanced mesh: baked geometry + per-instance model matrices and colours.

Reference: ``pybatchrender/renderer/node.py:12-342``.  Kept semantics:

* ``total_instances = num_scenes * instances_per_scene``; ``buf_instances`` is
  ``instances_per_scene`` for ``shared_across_scenes`` nodes, else ``total_instances`` (node.py:42-46)
* host mirrors ``transforms_b44`` / ``rot3_b33`` / ``scale_b11``; an "upload" recomputes
  ``transforms[:, :3, :3] = rot * scale`` and writes the column-packed matrices into ``matbuf``
  (node.py:110-126); ``lazy=True`` defers that upload to the next non-lazy setter (node.py:133,142,153)
* inputs of any leading shape flatten row-major to ``[B, k]`` -- instance index = scene * I + inst --
  and are truncated to ``buf_instances`` (node.py:131,139,151,159)
* constructor arguments ``positions/hprs/scales/colors`` are ignored with a warning (node.py:36-40);
  ``model_path=None`` makes a geometry-less node whose setters do nothing (node.py:73-77)

What changed: ``matbuf`` / ``colbuf`` are float32 tensors on the renderer's device ([B,16] and [B,4],
the exact byte layout of the reference's RGBA32F buffer textures) that the CUDA rasteriser reads in
place; nothing is serialised or re-uploaded per step.  On CUDA the upload is one fused kernel
(``pbr_pack_transforms``) instead of a chain of torch ops.

``set_pose`` (new) binds the node's matrices to *pose channels* -- position / HPR / scale given as
constants or as 1-D views of a device tensor (typically columns of the simulation state).  The
raster kernel then computes ``M = [Rz(h) Ry(p) Rx(r) * s | t]`` per instance itself when the frame
runs: the reference's per-step ``set_positions`` + ``set_hprs`` + upload (node.py:128-143) cost no
launch and no matrix traffic.  ``matbuf`` and the host mirrors are refreshed from the channels on
demand (reading them materialises the matrices with ``pbr_compose_transforms``, the same device
function the raster kernel uses); any generic setter ends the binding.
"""
from __future__ import annotations

import warnings
from collections.abc import Sequence
from typing import Literal

import torch

from .. import meshes as _meshes
from .shader_context import PBRShaderContext


class _NodeHandle:
    """Stand-in for the Panda3D ``NodePath`` the reference exposes as ``node.np``."""

    def __init__(self, owner: "PBRNode", name: str) -> None:
        self._owner = owner
        self._name = name
        self._removed = False

    def removeNode(self) -> None:
        self._removed = True
        self._owner._detach()

    remove_node = removeNode

    def isEmpty(self) -> bool:
        return self._removed

    is_empty = isEmpty

    def getName(self) -> str:
        return self._name


class PBRNode(PBRShaderContext):
    def __init__(self,
                 showbase,
                 model_path,
                 num_scenes: int = 1,
                 instances_per_scene: int = 1,
                 texture=None,
                 model_pivot_relative_point: tuple[float, float, float] | None = None,
                 model_scale: float | Sequence[float] | None = None,
                 model_hpr: Sequence[float] | None = None,
                 model_scale_units: Literal["relative", "absolute"] = "relative",
                 positions: torch.Tensor | None = None,
                 hprs: torch.Tensor | None = None,
                 scales: torch.Tensor | None = None,
                 colors: torch.Tensor | None = None,
                 backend: Literal["loop", "instanced"] = "instanced",
                 shared_across_scenes: bool = False,
                 parent=None,
                 name: str | None = None) -> None:
        if positions is not None or colors is not None or scales is not None or hprs is not None:
            warnings.warn("initializing positions, colors, scales, hprs through the constructor is not "
                          "implemented (as in the reference); they are ignored", stacklevel=3)
        super().__init__(showbase, backend=backend)
        self.num_scenes = int(num_scenes)
        self.instances_per_scene = int(instances_per_scene)
        self.shared_across = bool(shared_across_scenes)
        self.total_instances = self.num_scenes * self.instances_per_scene
        self.buf_instances = self.instances_per_scene if self.shared_across else self.total_instances
        self.model_pivot_relative_point = model_pivot_relative_point
        self.model_scale = model_scale
        self.model_hpr = model_hpr
        self.model_scale_units = model_scale_units
        self.shader_inputs: dict = {}
        self.name = name or (str(model_path) if isinstance(model_path, str) else "pbr_node")
        self.np = _NodeHandle(self, self.name)
        self._native_mesh = None          # device copy, created by the renderer on first use

        if model_path is not None and not (isinstance(model_path, str) and model_path == ""):
            self.mesh = _meshes.bake(_meshes.load_mesh(model_path), model_scale=model_scale,
                                     model_hpr=model_hpr, model_scale_units=model_scale_units,
                                     pivot_rel=model_pivot_relative_point)
            self.has_geometry = True
        else:
            self.mesh = None
            self.has_geometry = False

        self._register_self()
        self._attempt_camera_connect()
        self._attempt_light_connect()

        self.set_texture(texture)

        if self.has_geometry:
            dev, B = self.device, self.buf_instances
            self._matbuf = torch.zeros((B, 16), dtype=torch.float32, device=dev)
            self._pose = None              # dict(pos, hpr, scale) while the matrices are bound to pose channels
            self._pose_struct = None       # native pbr_pose_desc owned by the node (created by the first set_pose)
            self._pose_chans = None
            self._pose_cols = None
            self._mirror_stale = False
            self.colbuf = torch.ones((B, 4), dtype=torch.float32, device=dev)
            self._set_shader_input("instancesPerScene", self.instances_per_scene)
            self._set_shader_input("shareAcrossScenes", 1 if self.shared_across else 0)
            self._transforms_b44 = torch.eye(4, dtype=torch.float32, device=dev).repeat(B, 1, 1)
            self._rot3_b33 = torch.eye(3, dtype=torch.float32, device=dev).repeat(B, 1, 1)
            self._scale_b11 = torch.ones((B, 1, 1), dtype=torch.float32, device=dev)
            self._upload_current_transforms()

    # ------------------------------------------------------------------ pose binding + lazy mirrors
    def set_pose(self, pos=(0.0, 0.0, 0.0), hpr=(0.0, 0.0, 0.0), scale=1.0) -> None:
        """Bind the matrices to pose channels: each of ``pos[0..2]``, ``hpr[0..2]`` (radians, the
        reference's own convention R = Rz(H) Ry(P) Rx(R), shader_context.py:47-84) and ``scale`` is a
        float or a 1-D float32 view with one element per instance (any stride) on the node's device.
        The channels are read when a frame executes, so the tensors must hold the wanted state then
        (they are kept alive here).  Equivalent to ``set_positions`` + ``set_hprs`` + ``set_scales``."""
        if not self.has_geometry:
            return
        pos, hpr = tuple(pos), tuple(hpr)
        if len(pos) != 3 or len(hpr) != 3:
            raise ValueError("set_pose: pos and hpr take three channels each")
        native = getattr(self.base, "_native", None)
        if native is None or not self._matbuf.is_cuda:
            # no GPU: the generic torch path (host logic stays testable; rendering needs CUDA anyway)
            B = self.buf_instances

            def col(c):
                if isinstance(c, torch.Tensor):
                    return c.reshape(-1).to(device=self.device, dtype=torch.float32)
                return torch.full((B,), float(c), dtype=torch.float32, device=self.device)
            self.set_positions(torch.stack([col(c) for c in pos], 1), lazy=True)
            self.set_hprs(torch.stack([col(c) for c in hpr], 1), lazy=True)
            self.set_scales(col(scale).reshape(-1, 1))
            return
        # The node owns one native pose descriptor (pbr_pose_desc); a frame description points at it, so
        # re-binding channels here is all a step has to do.  Channels that did not change are skipped.
        chans = pos + hpr + (scale,)
        prev = self._pose_chans if self._pose_cols is None else None      # (columns were re-pointed since: rewrite all)
        if self._pose_struct is None:
            from .. import _native
            self._pose_struct = _native.new_pose_struct(self._matbuf)
        if self._pose is None:
            self._touch()                  # the frame description changes shape: posed from now on
        st, B = self._pose_struct, self.buf_instances
        for k in range(7):
            c = chans[k]
            if prev is not None and c is prev[k]:
                continue
            dst = st.pos[k] if k < 3 else (st.hpr[k - 3] if k < 6 else st.scale)
            if isinstance(c, torch.Tensor):
                if c.dim() != 1 or c.shape[0] != B or not c.is_cuda or c.dtype != torch.float32:
                    raise ValueError(f"set_pose: channel tensors must be 1-D float32 CUDA views of {B} elements")
                dst.ptr, dst.stride, dst.constant = c.data_ptr(), c.stride(0), 0.0
            else:
                if prev is not None and isinstance(prev[k], (int, float)) and float(prev[k]) == float(c):
                    continue
                dst.ptr, dst.stride, dst.constant = None, 0, float(c)
        self._pose_chans = chans           # keeps the channel tensors alive
        self._pose_cols = None
        self._pose = dict(pos=pos, hpr=hpr, scale=scale)
        self._mirror_stale = True

    def bind_pose_columns(self, state: torch.Tensor, columns) -> None:
        """Fast re-binding for renderers that feed columns of one ``[B, k]`` float32 state tensor every step
        (CartPole: ``((0, 0), (4, 2))`` = pos.x <- state[:, 0], hpr.P <- state[:, 2]): pairs of (channel index:
        0..2 position, 3..5 H/P/R, 6 scale; column).  Equivalent to ``set_pose`` with ``state[:, col]`` views for
        those channels and everything else unchanged, without creating the views.  Needs a previous ``set_pose``."""
        st = self._pose_struct
        if st is None or self._pose is None:
            raise RuntimeError("bind_pose_columns needs a pose bound with set_pose first")
        if isinstance(columns, dict):
            columns = tuple(columns.items())
        shape = state.shape
        if len(shape) != 2 or shape[0] != self.buf_instances or state.dtype != torch.float32 or not state.is_cuda:
            raise ValueError(f"bind_pose_columns: state must be a [{self.buf_instances}, k] float32 CUDA tensor")
        base, (s0, s1) = state.data_ptr(), state.stride()
        for k, col in columns:
            dst = st.pos[k] if k < 3 else (st.hpr[k - 3] if k < 6 else st.scale)
            dst.ptr, dst.stride = base + 4 * s1 * col, s0
        self._pose_cols = (state, columns)     # keeps the tensor alive; mirrors resolve channels through it
        self._mirror_stale = True

    def _pose_desc(self):
        """The node's native pose descriptor (None: matrices come from matbuf)."""
        return None if self._pose is None else self._pose_struct

    def _materialise_pose(self) -> None:
        if self._pose is not None:
            self.base._native.compose_structs([self._pose_struct], self.device)

    def _sync_mirrors(self) -> None:
        """Refresh transforms_b44 / rot3_b33 / scale_b11 from the bound pose (reference keeps them current on
        every setter, node.py:116-134; here they are derived when somebody looks)."""
        if not self._mirror_stale:
            return
        self._mirror_stale = False
        self._materialise_pose()
        B = self.buf_instances
        T = self._matbuf.view(B, 4, 4).transpose(1, 2).contiguous()
        sc = self._pose_chans[6]
        if self._pose_cols is not None:
            for k, col in self._pose_cols[1]:
                if k == 6:
                    sc = self._pose_cols[0][:, col]
        sc = sc.reshape(B, 1, 1).clone() if isinstance(sc, torch.Tensor) else torch.full(
            (B, 1, 1), float(sc), dtype=torch.float32, device=self.device)
        self._transforms_b44, self._scale_b11 = T, sc
        self._rot3_b33 = T[:, 0:3, 0:3] / sc

    def _end_pose(self) -> None:
        """A generic setter takes over: bring the mirrors up to date, then drop the binding."""
        if self.has_geometry and self._pose is not None:
            self._sync_mirrors()
            self._pose = None
            self._pose_chans = None
            self._pose_cols = None
            self._touch()

    @property
    def matbuf(self) -> torch.Tensor:
        self._materialise_pose()
        return self._matbuf

    @property
    def transforms_b44(self) -> torch.Tensor:
        self._sync_mirrors()
        return self._transforms_b44

    @transforms_b44.setter
    def transforms_b44(self, v: torch.Tensor) -> None:
        self._transforms_b44 = v

    @property
    def rot3_b33(self) -> torch.Tensor:
        self._sync_mirrors()
        return self._rot3_b33

    @rot3_b33.setter
    def rot3_b33(self, v: torch.Tensor) -> None:
        self._rot3_b33 = v

    @property
    def scale_b11(self) -> torch.Tensor:
        self._sync_mirrors()
        return self._scale_b11

    @scale_b11.setter
    def scale_b11(self, v: torch.Tensor) -> None:
        self._scale_b11 = v

    # ------------------------------------------------------------------ uploads
    def _as_rows(self, value, width: int) -> torch.Tensor:
        t = torch.as_tensor(value, dtype=torch.float32)
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        return t.reshape(-1, width)[: self.buf_instances]

    def _upload_mat(self, mats: torch.Tensor) -> None:
        if not self.has_geometry:
            return
        self._end_pose()
        self._matbuf.view(-1, 4, 4).copy_(mats.transpose(1, 2))
        self._touch()

    def _upload_current_transforms(self) -> None:
        if not self.has_geometry:
            return
        self._end_pose()
        B = self._matbuf.shape[0]
        if self.scale_b11.shape[0] != B:
            # the reference broadcasts a short scale tensor (rot3_b33 * scale_b11, node.py:119)
            if self.scale_b11.shape[0] != 1:
                raise ValueError(f"scale has {self.scale_b11.shape[0]} rows, node has {B} instances")
            self.scale_b11 = self.scale_b11.expand(B, 1, 1).contiguous()
        if self.transforms_b44.shape[0] != B or self.rot3_b33.shape[0] != B:
            raise ValueError(f"transforms / rotations have {self.transforms_b44.shape[0]} / {self.rot3_b33.shape[0]} "
                             f"rows, node has {B} instances")
        native = getattr(self.base, "_native", None)
        if native is not None and self._matbuf.is_cuda:
            native.pack_transforms(self.transforms_b44, self.rot3_b33, self.scale_b11, self._matbuf)
        else:
            self.transforms_b44[:, 0:3, 0:3] = self.rot3_b33 * self.scale_b11
            self._matbuf.view(-1, 4, 4).copy_(self.transforms_b44.transpose(1, 2))
        self._touch()

    def _touch(self) -> None:
        self._version = getattr(self, "_version", 0) + 1

    # ------------------------------------------------------------------ public setters
    def set_positions(self, pos_si3, lazy: bool = False) -> None:
        if not self.has_geometry:
            return
        self._end_pose()
        self.transforms_b44[:, 0:3, 3] = self._as_rows(pos_si3, 3)
        if not lazy:
            self._upload_current_transforms()

    def set_hprs(self, hpr_si3, lazy: bool = False) -> None:
        if not self.has_geometry:
            return
        self._end_pose()
        self.rot3_b33[:, :, :] = type(self)._rotation_mats_from_hpr(self._as_rows(hpr_si3, 3))
        if not lazy:
            self._upload_current_transforms()

    def set_scales(self, scale_si1, lazy: bool = False) -> None:
        if not self.has_geometry:
            return
        self._end_pose()
        if isinstance(scale_si1, (float, int)):
            self.scale_b11 = torch.full((self.buf_instances, 1, 1), float(scale_si1),
                                        dtype=torch.float32, device=self.device)
        else:
            self.scale_b11 = self._as_rows(scale_si1, 1).reshape(-1, 1, 1).clone()
        if not lazy:
            self._upload_current_transforms()

    def set_colors(self, col_si4) -> None:
        if not self.has_geometry:
            return
        self.colbuf.copy_(self._as_rows(col_si4, 4))
        self._touch()

    def set_transforms(self, mats_b44) -> None:
        """Full per-instance matrices; split into rotation and mean-column-norm scale (node.py:163-178)."""
        if not self.has_geometry:
            return
        self._end_pose()
        m = torch.as_tensor(mats_b44, dtype=torch.float32).to(self.device).reshape(-1, 4, 4)
        if m.shape[0] != self.buf_instances:
            raise ValueError(f"set_transforms: got {m.shape[0]} matrices for {self.buf_instances} instances")
        self.transforms_b44 = m.clone()
        r = self.transforms_b44[:, 0:3, 0:3].clone()
        s = torch.linalg.norm(r, dim=2).mean(dim=1).clamp_min(1e-8)
        self.scale_b11 = s.reshape(-1, 1, 1)
        self.rot3_b33 = r / self.scale_b11
        self._upload_current_transforms()

    # ------------------------------------------------------------------ misc API
    def set_texture(self, texture=None) -> None:
        """``None`` / ``False``: untextured.  A file name, ``[h, w, 3|4]`` uint8 array / tensor (row 0
        = top of the picture) or PIL image: sampled with the mesh's UVs (GL_REPEAT, GL_LINEAR) and
        multiplied into the instance colour (reference node.py:277-287, basic.frag:31-37).  ``True``
        without an image keeps the node white, like Panda3D's default texture."""
        self.texture_image = None            # numpy [h, w, 4] uint8, row 0 = v 0 (bottom row)
        self._native_texture = None
        if texture is None or texture is False:
            self._set_shader_input("useTexture", 0.0)
        else:
            if texture is not True:
                import numpy as np
                if isinstance(texture, (str, bytes)) or hasattr(texture, "__fspath__"):
                    from PIL import Image
                    texture = Image.open(texture)
                if hasattr(texture, "convert"):                      # PIL image
                    texture = np.asarray(texture.convert("RGBA"))
                if isinstance(texture, torch.Tensor):
                    texture = texture.detach().cpu().numpy()
                img = np.asarray(texture)
                if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] not in (3, 4):
                    raise ValueError("texture must be a file, a PIL image or a [h, w, 3|4] uint8 array")
                if img.shape[2] == 3:
                    img = np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2)
                self.texture_image = np.ascontiguousarray(img[::-1])
            self._set_shader_input("useTexture", 1.0)
        self._touch()

    def reparent_to(self, parent) -> None:
        raise NotImplementedError("PBRNode.reparent_to is not implemented yet")

    def pivot_to_rel(self, relative_point: tuple[float, float, float] | None = None) -> None:
        """Re-centre the baked geometry so the bounds-relative point becomes the origin (node.py:313-341)."""
        if relative_point is None or not self.has_geometry:
            return
        _meshes.bake(self.mesh, pivot_rel=relative_point)
        self._native_mesh = None
        self._touch()

    def _set_lighting_strength(self, strength: float, overwrite: bool = False) -> None:
        if overwrite or not hasattr(self.base, "_pbr_light"):
            self._set_shader_input("lightingStrength", float(strength))

    def _set_lighting(self, dir_dir, dir_col, amb_col, overwrite: bool = False) -> None:
        if overwrite or not hasattr(self.base, "_pbr_light"):
            self._set_shader_input("dirLightDir", tuple(float(x) for x in dir_dir))
            self._set_shader_input("dirLightCol", tuple(float(x) for x in dir_col))
            self._set_shader_input("ambientCol", tuple(float(x) for x in amb_col))

    def _register_self(self) -> None:
        if not hasattr(self.base, "_pbr_nodes"):
            self.base._pbr_nodes = []
        if self not in self.base._pbr_nodes:
            self.base._pbr_nodes.append(self)
            if hasattr(self.base, "_scene_changed"):
                self.base._scene_changed()

    def _detach(self) -> None:
        """``node.np.removeNode()``: stop drawing this node (Steering rebuilds obstacles this way)."""
        nodes = getattr(self.base, "_pbr_nodes", [])
        if self in nodes:
            nodes.remove(self)
            if hasattr(self.base, "_scene_changed"):
                self.base._scene_changed()

    def _attempt_camera_connect(self) -> None:
        cam = getattr(self.base, "_pbr_cam", None)
        if cam:
            cam.attach(self)

    def _attempt_light_connect(self) -> None:
        light = getattr(self.base, "_pbr_light", None)
        if light:
            light.attach(self)
