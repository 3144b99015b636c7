"""A frame *stream* through the matching path: every frame's feature maps are uploaded once and
matched against the previous frame's, which stayed on the device -- the residency
depth_estimation_api.lua keeps in `last_filtered` (:68-72, :188-190), batched.  With independent
pairs both frames cross PCIe for every pair; in a stream only the new one does, which moves the
end-to-end bound from the PCIe link to the kernels.

    s = FeatureStream(maxh, maxw, channels, h, w, batch=16)
    s.prime(first_frame_features)                  # [C,H,W] host (pinned) or device
    for frames in batches:                         # [B,C,H,W] pinned host memory
        res = s.push(frames)                       # dict of pinned host arrays, valid after res.wait()

Pair i of a batch is (frame i-1, frame i): in1 = prepareInput's crop of the older frame, in2 =
the newer one, both strided views of one device buffer (no copies).  H2D of batch k+1 (copy
stream), kernels of batch k (compute stream) and D2H of batch k-1 (result stream) overlap.
"""
import math

import numpy as np

from . import api

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


class _Pending:
    def __init__(self, host, event):
        self.host, self._event = host, event

    def wait(self):
        self._event.synchronize()
        return self.host

    def __getitem__(self, k):
        return self.host[k]


class FeatureStream:
    def __init__(self, maxh, maxw, channels, h, w, batch=16, want=("index", "pmax", "score_thr"),
                 canvas=True, device=0, ctx=None):
        if torch is None or not torch.cuda.is_available():
            raise api.DepthMatchError(api._lib.DM_ERR_CUDA, "FeatureStream needs a CUDA device (no CPU fallback)")
        self.maxh, self.maxw, self.C, self.H, self.W, self.B = maxh, maxw, channels, h, w, batch
        self.h1, self.w1 = h - maxh + 1, w - maxw + 1
        self.oy, self.ox = math.ceil(maxh / 2) - 1, math.ceil(maxw / 2) - 1
        self.want, self.canvas = tuple(want), (h, w) if canvas else None
        self.dev = torch.device("cuda", device)
        self.ctx = ctx or api.Context(device)
        # two device buffers of batch + 1 frames: slot 0 holds the last frame of the previous batch
        self.buf = [torch.empty((batch + 1, channels, h, w), device=self.dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.compute_stream = torch.cuda.Stream(self.dev)
        self.result_stream = torch.cuda.Stream(self.dev)
        self.uploaded = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        shapes = {"index": ((batch, self.h1, self.w1), torch.int64), "min_ssd": ((batch, self.h1, self.w1), torch.float32),
                  "pmax": ((batch, self.h1, self.w1), torch.float32),
                  "index_thr": ((batch, self.h1, self.w1), torch.int64),
                  "score_thr": ((batch, self.h1, self.w1), torch.float32),
                  "soft_yx": ((batch, 2, self.h1, self.w1), torch.float32),
                  "flow_full": ((batch, 2, h, w), torch.float32)}
        names = list(self.want) + (["flow_full"] if canvas else [])
        self.out_dev = [{k: torch.empty(shapes[k][0], dtype=shapes[k][1], device=self.dev) for k in names}
                        for _ in range(2)]
        self.out_host = [{k: torch.empty(shapes[k][0], dtype=shapes[k][1]).pin_memory() for k in names}
                         for _ in range(2)]
        self.results_done = [torch.cuda.Event() for _ in range(2)]
        self.k = 0
        self.primed = False

    def prime(self, frame):
        """The first frame of the stream: uploaded, nothing to match yet."""
        t = frame if api._is_torch(frame) else torch.from_numpy(np.ascontiguousarray(frame, np.float32))
        with torch.cuda.stream(self.copy_stream):
            self.buf[0][0].copy_(t, non_blocking=True)
            self.uploaded[0].record(self.copy_stream)
        self.carry_ready = self.uploaded[0]
        self.primed = True
        self.k = 0

    def push(self, frames):
        """frames: [B,C,H,W] float32 in pinned host memory (or on the device).  Returns a handle
        whose .wait() gives {name: pinned host array} for the B pairs (previous, frame 0),
        (frame 0, frame 1), ...; the arrays are reused two pushes later."""
        if not self.primed:
            raise api.DepthMatchError(api._lib.DM_ERR_INVALID, "FeatureStream.push before prime()")
        t = frames if api._is_torch(frames) else torch.from_numpy(frames)
        if tuple(t.shape) != (self.B, self.C, self.H, self.W) or t.dtype != torch.float32:
            raise api.DepthMatchError(api._lib.DM_ERR_INVALID, "FeatureStream.push: frames must be float32 %s"
                                      % ((self.B, self.C, self.H, self.W),))
        cur, nxt = self.k & 1, (self.k + 1) & 1
        buf = self.buf[cur]
        # upload behind the kernels that last read this buffer
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[cur]) if self.k >= 2 else None
            buf[1:].copy_(t, non_blocking=True)
            self.uploaded[cur].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.uploaded[cur])
            self.compute_stream.wait_event(self.carry_ready)
            if self.k >= 2:
                self.compute_stream.wait_event(self.results_done[cur])   # out_dev[cur] was copied out
            in1 = buf[:self.B, :, self.oy:self.oy + self.h1, self.ox:self.ox + self.w1]
            in2 = buf[1:]
            api.match_extract(in1, in2, self.maxh, self.maxw, canvas=self.canvas, want=self.want, ctx=self.ctx,
                              out=self.out_dev[cur])
            # the newest frame opens the next batch
            self.buf[nxt][0].copy_(buf[self.B], non_blocking=True)
            carry = torch.cuda.Event()
            carry.record(self.compute_stream)
            self.carry_ready = carry
            self.consumed[cur].record(self.compute_stream)
        with torch.cuda.stream(self.result_stream):
            self.result_stream.wait_event(self.consumed[cur])
            for name, d in self.out_dev[cur].items():
                self.out_host[cur][name].copy_(d, non_blocking=True)
            self.results_done[cur].record(self.result_stream)
        self.k += 1
        return _Pending({n: v.numpy() for n, v in self.out_host[cur].items()}, self.results_done[cur])

    def synchronize(self):
        for s in (self.copy_stream, self.compute_stream, self.result_stream):
            s.synchronize()
