"""``PBRCam`` -- per-scene view matrices, one shared OpenGL-style projection, packed VP buffer.

Reference: ``pybatchrender/renderer/camera.py:10-434``.  Kept semantics:

* tile table ``(u0,u1,v0,v1) = (c/cols, (c+1)/cols, 1-(r+1)/rows, 1-r/rows)``, ``c = i % cols``,
  ``r = i // cols``                                                        (camera.py:110-135)
* projection: ``f = 1/tan(fov_y/2)``, ``P00 = f/aspect``, ``P11 = f``, ``P22 = (zf+zn)/(zn-zf)``,
  ``P23 = 2 zf zn/(zn-zf)``, ``P32 = -1`` with **aspect = window X / window Y** (quirk Q1,
  camera.py:150-166); cached when ``fixed_projection``
* look-at basis ``f = norm(fwd)``, ``s = norm(f x up)``, ``u = s x f``; rows of V are ``s, u, -f``
  and the translation column is ``(-s.e, -u.e, +f.e)``; a missing ``up`` reuses the current
  ``V[:,1,:3]``, a missing ``forward`` reuses ``-V[:,2,:3]``            (camera.py:201-240)
* ``VP = P @ V`` per scene, stored column-packed: texel j of scene k = column j (camera.py:252-262)
* ``(3,)`` inputs broadcast to all scenes (camera.py:183-193); ``set_positions(keep_lookat=True)``
  raises (camera.py:275-279)

Deviation (documented, quirk Q5): ``set_projection`` really recomputes P and VP -- in the reference it
only clears a cache and re-uploads the old VP (camera.py:168-181).

``viewbuf`` is a float32 ``[K,16]`` tensor on the renderer's device, read in place by the kernels.
The tile table is kept (``tilebuf``) for API parity; the rasteriser renders each scene directly
into ``out[scene]`` so it only needs the tile *size*.
"""
from __future__ import annotations

import math
from typing import Literal

import torch

from .shader_context import PBRShaderContext


class PBRCam(PBRShaderContext):
    def __init__(self, showbase,
                 num_scenes: int,
                 cols: int | None = None,
                 rows: int | None = None,
                 backend: Literal["loop", "instanced"] = "instanced",
                 fov_y_deg: float = 55.0,
                 z_near: float = 0.05,
                 z_far: float = 100.0,
                 auto_tiles: bool = True,
                 fixed_projection: bool = True) -> None:
        super().__init__(showbase, backend=backend)
        self.num_scenes = int(num_scenes)
        self.fov_y_deg = float(fov_y_deg)
        self.z_near = float(z_near)
        self.z_far = float(z_far)
        self.fixed_projection = fixed_projection
        dev, K = self.device, self.num_scenes

        self.viewbuf = torch.zeros((max(1, K), 16), dtype=torch.float32, device=dev)
        self.tilebuf = torch.zeros((max(1, K), 4), dtype=torch.float32, device=dev)
        self.cols, self.rows = cols, rows
        self._set_tiles()

        self.eye_k3 = torch.zeros((K, 3), dtype=torch.float32, device=dev)
        self.forward_k3 = torch.zeros((K, 3), dtype=torch.float32, device=dev)
        self.up_k3 = torch.zeros((K, 3), dtype=torch.float32, device=dev)
        self.hpr_k3 = torch.zeros((K, 3), dtype=torch.float32, device=dev)
        self.target_k3 = torch.zeros((K, 3), dtype=torch.float32, device=dev)
        self.V_k44 = torch.eye(4, dtype=torch.float32, device=dev).repeat(K, 1, 1)
        self.VP_k44 = torch.zeros((K, 4, 4), dtype=torch.float32, device=dev)

        self._proj_cache_key = None
        self._proj_cache = None
        self.P_k44 = self._get_projection()

        # True while every scene provably has the same view (every input that feeds V -- given or reused from
        # an earlier call -- was a broadcast (3,) vector): lets the renderer pre-render shared nodes once
        self._uni = {"eye": True, "fwd": True, "up": True}
        self._update_view(eye_k3=torch.tensor([0.0, -12.0, 0.0]),
                          forward_k3=torch.tensor([0.0, 1.0, 0.0]),
                          up_k3=torch.tensor([0.0, 0.0, 1.0]))
        self._update_vp()
        self.sync_from_base_cam = False
        if getattr(self.base, "_pbr_nodes", None):
            self.attach_all()
        self._register_self()

    # ------------------------------------------------------------------ shader-input plumbing
    def attach(self, node) -> None:
        node._set_shader_input("K", self.num_scenes)
        node._auto_screen_size_input()

    def attach_many(self, nodes) -> None:
        for n in nodes:
            self.attach(n)

    def attach_all(self) -> None:
        self._set_shader_input("K", self.num_scenes)
        self._auto_screen_size_input()

    # ------------------------------------------------------------------ tiles
    def _set_tiles(self) -> None:
        if self.num_scenes is None and (self.cols is None or self.rows is None):
            raise ValueError("num_scenes or (cols and rows) must be provided")
        if self.cols is None and self.rows is None:
            self.cols = math.ceil(math.sqrt(self.num_scenes))
        if self.rows is None:
            self.rows = math.ceil(self.num_scenes / self.cols)
        if self.num_scenes is None:
            self.num_scenes = self.cols * self.rows
        K = int(self.num_scenes)
        i = torch.arange(K, dtype=torch.float32)
        cols, rows = float(self.cols), float(self.rows)
        c = torch.remainder(i, cols)
        r = torch.floor(i / cols)
        tiles = torch.stack([c / cols, (c + 1.0) / cols, 1.0 - (r + 1.0) / rows, 1.0 - r / rows], dim=1)
        self.tilebuf[:K].copy_(tiles.to(torch.float32))

    def _set_tiles_from_array(self, tiles_k4: torch.Tensor) -> None:
        self.rows = tiles_k4.shape[0]
        self.cols = tiles_k4.shape[1]
        self.num_scenes = self.cols * self.rows
        self.tilebuf = tiles_k4.to(torch.float32).reshape(-1, 4).contiguous().to(self.device)

    # ------------------------------------------------------------------ projection
    def _get_projection(self) -> torch.Tensor:
        if self.fixed_projection and self._proj_cache is not None:
            return self._proj_cache
        win = self.base.win
        aspect = max(1e-6, float(win.getXSize()) / max(1, win.getYSize()))
        key = (self.fov_y_deg, self.z_near, self.z_far, float(aspect))
        if self._proj_cache_key != key or self._proj_cache is None:
            f = 1.0 / math.tan(math.radians(self.fov_y_deg) * 0.5)
            zn, zf = self.z_near, self.z_far
            P = torch.zeros((4, 4), dtype=torch.float32)
            P[0, 0] = f / aspect
            P[1, 1] = f
            P[2, 2] = (zf + zn) / (zn - zf)
            P[2, 3] = (2.0 * zf * zn) / (zn - zf)
            P[3, 2] = -1.0
            self._proj_cache_key = key
            self._proj_cache = P.to(self.device)
        return self._proj_cache

    def set_projection(self, fov_y_deg: float | None = None, z_near: float | None = None,
                       z_far: float | None = None) -> None:
        if fov_y_deg is not None:
            self.fov_y_deg = float(fov_y_deg)
        if z_near is not None:
            self.z_near = float(z_near)
        if z_far is not None:
            self.z_far = float(z_far)
        self._proj_cache_key = None
        self._proj_cache = None
        self.P_k44 = self._get_projection()
        self._update_vp()

    # ------------------------------------------------------------------ view
    def _ensure_kx3(self, arr, name: str) -> torch.Tensor:
        a = torch.as_tensor(arr, dtype=torch.float32)
        if a.device != self.device:
            a = a.to(self.device, non_blocking=True)
        if a.ndim == 1:
            a = a.unsqueeze(0)
        if a.shape[-1] != 3:
            raise ValueError(f"{name} must have shape (K,3) or (3,), got {tuple(a.shape)}")
        if a.shape[0] == 1 and self.num_scenes > 1:
            a = a.repeat(self.num_scenes, 1)
        if a.shape[0] != self.num_scenes:
            raise ValueError(f"{name} first dim must be K={self.num_scenes}, got {a.shape[0]}")
        return a

    @staticmethod
    def _normalize(v: torch.Tensor) -> torch.Tensor:
        return v / torch.linalg.norm(v, dim=-1, keepdim=True).clamp_min(1e-8)

    def _update_view(self, eye_k3=None, forward_k3=None, up_k3=None, fwd_uniform=None, up_uniform=None) -> None:
        """``fwd_uniform`` / ``up_uniform``: whether the given forward / up input is the same for every
        scene.  An omitted input reuses the current rows of V (reference camera.py:221), which may be
        per-scene from an earlier call, so the uniformity of the *result* is tracked here: the basis
        (rows s, u) is uniform only when both of its inputs -- given or reused -- are."""
        basis_changed = forward_k3 is not None or up_k3 is not None
        if basis_changed:
            f_uni = self._uni["fwd"] if forward_k3 is None else bool(self._bcast(forward_k3) if fwd_uniform is None else fwd_uniform)
            u_uni = self._uni["up"] if up_k3 is None else bool(self._bcast(up_k3) if up_uniform is None else up_uniform)
            self._uni["fwd"] = f_uni
            self._uni["up"] = f_uni and u_uni
        if eye_k3 is not None:
            self._uni["eye"] = self._bcast(eye_k3)
        eye_changed = eye_k3 is not None
        if not basis_changed and not eye_changed:
            return
        V = self.V_k44
        if basis_changed:
            if forward_k3 is not None:
                f = self._normalize(self._ensure_kx3(forward_k3, "forward_k3"))
            else:
                f = -V[:, 2, 0:3]
            if up_k3 is not None:
                up_in = self._normalize(self._ensure_kx3(up_k3, "up_k3"))
            else:
                up_in = V[:, 1, 0:3]
            s = self._normalize(torch.cross(f, up_in, dim=-1))
            u = torch.cross(s, f, dim=-1)
            V[:, 0, 0:3] = s
            V[:, 1, 0:3] = u
            V[:, 2, 0:3] = -f
        else:
            s, u, f = V[:, 0, 0:3], V[:, 1, 0:3], -V[:, 2, 0:3]
        if eye_changed:
            self.eye_k3 = self._ensure_kx3(eye_k3, "eye_k3")
        e = self.eye_k3
        V[:, 0, 3] = -(s * e).sum(-1)
        V[:, 1, 3] = -(u * e).sum(-1)
        V[:, 2, 3] = (f * e).sum(-1)

    @staticmethod
    def _fwd_up_from_hpr(hpr_k3: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        R = PBRShaderContext._rotation_mats_from_hpr(hpr_k3)
        return R[:, :, 1].clone(), R[:, :, 2].clone()     # R @ (0,1,0), R @ (0,0,1)

    def _update_vp(self) -> None:
        self.VP_k44 = torch.matmul(self.P_k44, self.V_k44).reshape(-1, 4, 4).to(torch.float32)
        self._upload_viewproj(self.VP_k44)

    def _upload_viewproj(self, VP_k44: torch.Tensor) -> None:
        self.viewbuf[: VP_k44.shape[0]].view(-1, 4, 4).copy_(VP_k44.transpose(1, 2))
        self._version = getattr(self, "_version", 0) + 1

    def get_vp(self) -> torch.Tensor:
        return self.VP_k44.clone()

    @property
    def uniform(self) -> bool:
        """All scenes share one VP (known from how the camera was set, never from a device read)."""
        return self.num_scenes == 1 or all(self._uni.values())

    def _bcast(self, x) -> bool:
        t = torch.as_tensor(x)
        return self.num_scenes == 1 or t.ndim == 1 or t.shape[0] == 1

    # ------------------------------------------------------------------ public API
    def look_at(self, target_k3, lazy: bool = False) -> None:
        self._update_view(forward_k3=self._fwd_from_lookat(target_k3),
                          fwd_uniform=self._uni["eye"] and self._bcast(target_k3))
        if not lazy:
            self._update_vp()

    def set_positions(self, eye_k3, keep_lookat: bool = False, lazy: bool = False) -> None:
        if keep_lookat:
            raise NotImplementedError(
                "Target can't be reconstructed from eye and forward, needs to be stored in memory then")
        self.set_eye(eye_k3, lazy=lazy)

    def set_positions_and_lookat(self, eye_k3, target_k3, lazy: bool = False) -> None:
        uni = self._bcast(eye_k3) and self._bcast(target_k3)
        eye = self._ensure_kx3(eye_k3, "eye_k3")
        target = self._ensure_kx3(target_k3, "target_k3")
        self._update_view(eye_k3=eye, forward_k3=self._fwd_from_lookat(target, eye), fwd_uniform=uni)
        self._uni["eye"] = self._bcast(eye_k3)
        if not lazy:
            self._update_vp()

    def _fwd_from_lookat(self, target_k3, eye_k3=None) -> torch.Tensor:
        eye = self._ensure_kx3(eye_k3 if eye_k3 is not None else self.eye_k3, "eye_k3")
        return self._ensure_kx3(target_k3, "target_k3") - eye

    def set_hprs(self, hpr_k3, lazy: bool = False) -> None:
        uni = self._bcast(hpr_k3)
        fwd, up = self._fwd_up_from_hpr(self._ensure_kx3(hpr_k3, "hpr_k3"))
        self._update_view(forward_k3=fwd, up_k3=up, fwd_uniform=uni, up_uniform=uni)
        if not lazy:
            self._update_vp()

    def set_eye(self, eye_k3, lazy: bool = False) -> None:
        self._update_view(eye_k3=eye_k3)
        if not lazy:
            self._update_vp()

    def set_forward(self, forward_k3, lazy: bool = False) -> None:
        self._update_view(forward_k3=forward_k3)
        if not lazy:
            self._update_vp()

    def set_up(self, up_k3, lazy: bool = False) -> None:
        self._update_view(up_k3=up_k3)
        if not lazy:
            self._update_vp()

    def set_right(self, right_k3, lazy: bool = False) -> None:
        right = self._normalize(self._ensure_kx3(right_k3, "right_k3"))
        f = -self.V_k44[:, 2, 0:3]
        self._update_view(forward_k3=f, up_k3=torch.cross(right, f, dim=-1), fwd_uniform=self._uni["fwd"],
                          up_uniform=self._uni["fwd"] and self._bcast(right_k3))
        if not lazy:
            self._update_vp()

    def get_eye(self) -> torch.Tensor:
        return self.eye_k3.clone()

    def get_forward(self) -> torch.Tensor:
        return (-self.V_k44[:, 2, 0:3]).clone()

    def get_up(self) -> torch.Tensor:
        return self.V_k44[:, 1, 0:3].clone()

    def get_right(self) -> torch.Tensor:
        return self.V_k44[:, 0, 0:3].clone()

    def enable_base_cam_sync(self) -> None:      # onscreen-only feature of the reference: no-op
        self.sync_from_base_cam = False

    def disable_base_cam_sync(self) -> None:
        self.sync_from_base_cam = False

    def start_tasks(self, taskMgr=None, name: str = "pbr_cam_update") -> None:
        return None

    def _register_self(self) -> None:
        self.base._pbr_cam = self
