"""Frame grabbers -- kept for API parity only.

Reference: ``pybatchrender/renderer/frame_grabber.py:22-170``.  There the grabbers are the readback
stage: ``GPUFrameGrabber`` maps the GL colour texture into CUDA and copies it (D2D) into a CuPy
buffer, ``CPUFrameGrabber`` reads it back to RAM; both return the whole tiled window, flipped so that
row 0 is the top.  Here nothing is read back -- the raster kernel writes every scene straight into
the caller's tensor -- so a grabber is a thin view: ``grab()`` asks the renderer for the window image
(``PBRRenderer.grab_pixels``) on the device (GPU grabber) or copies it to the host (CPU grabber).
Code that constructed a grabber by hand (``GPUFrameGrabber(renderer, renderer.offscreen_tex)``) keeps
working; failures raise instead of yielding black frames.
"""
from __future__ import annotations

import torch


class BaseFrameGrabber:
    def __init__(self, base, tex=None, readonly: bool = True) -> None:
        self.base = base
        self.tex = tex

    def grab(self) -> torch.Tensor:          # pragma: no cover - abstract
        raise NotImplementedError

    def close(self) -> None:
        return None


class GPUFrameGrabber(BaseFrameGrabber):
    """``grab()`` -> uint8 CUDA tensor ``[rows*H, cols*W, C]`` (row 0 = top of the window)."""

    def grab(self) -> torch.Tensor:
        img = self.base.grab_pixels()
        if not img.is_cuda:
            raise RuntimeError("GPUFrameGrabber needs a renderer on a CUDA device")
        return img


class CPUFrameGrabber(BaseFrameGrabber):
    """``grab()`` -> the same image in host memory (device -> host copy of the finished frame)."""

    def grab(self) -> torch.Tensor:
        return self.base.grab_pixels().cpu()
