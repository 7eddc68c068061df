"""Frame-render pipeline around the fused kernel (replaces the per-frame host work of
renderer.py:58-77,126-178).

The reference renders a frame, synchronises, pulls rgb / depth to the host as fp32 (`.cpu()`), converts
them with numpy / cv2 / matplotlib and only then starts the next frame.  Here a frame is

    one fused launch per active field  ->  one post-processing kernel (8-bit BGR + colour-mapped depth)

whose outputs are written by the GPU straight into pinned host buffers (zero-copy, 6 bytes per pixel instead
of 16), with a CUDA event per frame.  `submit()` never blocks; `fetch()` returns the OLDEST finished frame,
so encoding / writing frame i on the host overlaps the rendering of frames i+1..i+depth-1.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .tensorf import _ptr, _stream


def jet_lut():
    """cv2.COLORMAP_JET as a [256,3] uint8 BGR table (what visualize_depth applies, utils/utils.py:195)."""
    import cv2
    return cv2.applyColorMap(np.arange(256, dtype=np.uint8)[:, None], cv2.COLORMAP_JET)[:, 0, :].copy()


class FramePipeline:
    def __init__(self, local_tensorfs, W, H, depth_minmax=(0.0, 5.0), n_buffers=3, lut=None, keep_float=False):
        self.lt, self.W, self.H = local_tensorfs, int(W), int(H)
        self.dev = local_tensorfs.blending_weights.device
        if self.dev.type != "cuda":
            raise RuntimeError("FramePipeline needs the scene model on a CUDA device")
        self.n = self.W * self.H
        self.minmax = (float(depth_minmax[0]), float(depth_minmax[1]))
        lut = jet_lut() if lut is None else np.ascontiguousarray(lut, np.uint8).reshape(256, 3)
        self.lut = torch.from_numpy(lut).to(self.dev)
        self.ids = torch.arange(self.n, dtype=torch.int64, device=self.dev)
        self.keep_float = keep_float
        self.slots = []
        for _ in range(int(n_buffers)):
            s = {"rgb8": torch.empty(self.n, 3, dtype=torch.uint8).pin_memory(),
                 "depth8": torch.empty(self.n, 3, dtype=torch.uint8).pin_memory(),
                 "event": torch.cuda.Event(), "busy": False, "tag": None}
            if keep_float:
                s["rgb"] = torch.empty(self.n, 3).pin_memory()
                s["depth"] = torch.empty(self.n).pin_memory()
            self.slots.append(s)
        self.head = self.tail = 0                       # ring: tail = oldest in flight, head = next to fill

    def in_flight(self):
        return sum(s["busy"] for s in self.slots)

    @torch.no_grad()
    def submit(self, view_ids, tag=None, **forward_kw):
        """Enqueues one frame (all W*H rays of `view_ids`; kwargs as LocalTensorfs.forward: cam2world,
        world2rf, blending_weights, test_id, floater_thresh ...).  Raises if every buffer is in flight."""
        s = self.slots[self.head]
        if s["busy"]:
            raise RuntimeError("all frame buffers are in flight: fetch() before submitting more")
        out = (s["rgb"], s["depth"]) if self.keep_float else None
        rgb, depth, _, _ = self.lt(self.ids, view_ids, self.W, self.H, is_train=False, out=out, **forward_kw)
        with torch.cuda.device(self.dev):
            st = _stream(self.dev)
            _lib.check(_lib.lib().lrf_frame_to_u8(_ptr(rgb), 3, _ptr(depth), 1, self.n, self.minmax[0],
                                                  self.minmax[1], _ptr(self.lut), _ptr(s["rgb8"]),
                                                  _ptr(s["depth8"]), st))
            s["event"].record(torch.cuda.current_stream(self.dev))
        s["busy"], s["tag"] = True, tag
        s["hold"] = (rgb, depth)                        # keep device outputs alive until the event has passed
        self.head = (self.head + 1) % len(self.slots)

    def fetch(self):
        """Blocks until the oldest submitted frame is complete -> dict(tag, rgb8 [H,W,3] uint8 BGR,
        depth8 [H,W,3] uint8 BGR, and rgb [H,W,3] / depth [H,W] float32 when keep_float).  The arrays are
        views of the pinned buffer: valid until `len(buffers)` more frames have been submitted."""
        s = self.slots[self.tail]
        if not s["busy"]:
            return None
        s["event"].synchronize()
        s["busy"] = False
        s.pop("hold", None)
        self.tail = (self.tail + 1) % len(self.slots)
        out = {"tag": s["tag"], "rgb8": s["rgb8"].numpy().reshape(self.H, self.W, 3),
               "depth8": s["depth8"].numpy().reshape(self.H, self.W, 3)}
        if self.keep_float:
            out["rgb"] = s["rgb"].numpy().reshape(self.H, self.W, 3)
            out["depth"] = s["depth"].numpy().reshape(self.H, self.W)
        return out

    def render(self, frames, **forward_kw):
        """Generator over `frames` (iterable of view-id tensors, or (view_ids, extra kwargs) pairs): keeps
        the ring full and yields finished frames in order."""
        for f in frames:
            view, extra = (f if isinstance(f, tuple) else (f, {}))
            if self.in_flight() == len(self.slots):
                yield self.fetch()
            self.submit(view, **{**forward_kw, **extra})
        while self.in_flight():
            yield self.fetch()
