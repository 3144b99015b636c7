"""Host-side mirror of the reference's Lua surface (see package docstring).

Arrays may be numpy arrays (host memory: staged by the library, call complete at
return) or torch CUDA tensors (device pointers: enqueued on torch's current
stream).  Outputs follow the inputs' kind.  Index conventions are the
reference's (1-based window indices, LongTensor = int64).
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import (DM_FLAG_ASYNC, DM_FLAG_DIFF_SSD, DM_FLAG_EXACT_SSD, DM_FLAG_TIE_MIDDLE, DM_VOLUME_EXACT, DM_VOLUME_NEG_SOFTMAX,
                   DM_VOLUME_SSD, DepthMatchError, check, dm_extract_out, dm_pair)

__all__ = [
    "Context", "default_context", "Geometry", "nn", "extractoutput", "yx2x", "x2yx",
    "centered2onebased", "onebased2centered", "getMiddleIndex", "prepareInput", "getModel",
    "DenseMatch", "processOutput", "getOutputConfidences", "getOutputConfidences2", "yx2xMulti",
    "x2yxMulti", "x2yxMulti2", "x2yxMultiNumber", "getModelMultiscale", "multiscaleLength",
    "getRMax", "getC2PMask", "getP2CMask", "cartesian2polar", "polar2cartesian", "getKOutput",
    "getP2CMaskOF", "flow2depth", "match_extract", "match_extract_raw_ssd", "match_volume", "round_lua",
    "postProcessImage", "enlargeMask", "radial", "computeDepthMapFromFlow",
    "Filter", "getFilter", "getRadialFilter", "getMultiscalePrefilter", "downsample", "multiscaleInputs",
]

try:  # torch is plumbing only (device memory + streams); the package works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


# ------------------------------------------------------------------ context
class Context:
    """One dm_ctx: a device, a stream, a workspace.  One per GPU per thread."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.dm_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._stream = "own"

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        check(self._lib.dm_synchronize(self._h))

    def launch_count(self):
        return int(self._lib.dm_launch_count(self._h))

    def set_profiling(self, on=True):
        check(self._lib.dm_set_profiling(self._h, 1 if on else 0))

    def last_kernel_ms(self):
        ms = C.c_float(0)
        check(self._lib.dm_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def set_option(self, name, value):
        """Tuning / diagnostic switch (dm_set_option): e.g. ("ssd_form", "dot" | "diff" | "auto")."""
        check(self._lib.dm_set_option(self._h, str(name).encode(), str(value).encode()))

    def last_counts(self):
        """(pixels handed to the entry-by-entry rescore, pixels through the exact thresholded pass)
        of the most recent match_extract on this context; synchronises."""
        a, b = C.c_int64(-1), C.c_int64(-1)
        check(self._lib.dm_last_counts(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def use_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (torch.cuda.current_stream().cuda_stream; 0 is the
        legacy default stream) or "own" for the context's private stream."""
        if cuda_stream != self._stream:
            if cuda_stream == "own":
                check(self._lib.dm_reset_stream(self._h))
            else:
                check(self._lib.dm_set_stream(self._h, C.c_void_p(int(cuda_stream))))
            self._stream = cuda_stream


_default = {}


def default_context(device=0):
    ctx = _default.get(device)
    if ctx is None:
        ctx = _default[device] = Context(device)
    return ctx


class _Args:
    """Collects pointers for one library call and remembers what must stay alive."""

    def __init__(self, ctx=None):
        self.keep = []
        self.on_device = False
        self.ctx = ctx

    def inp(self, x, dtype=np.float32):
        if _is_torch(x):
            tdt = {np.float32: torch.float32, np.int64: torch.int64}[dtype]
            if x.dtype != tdt or not x.is_contiguous():
                x = x.to(tdt).contiguous()
            self.keep.append(x)
            if x.is_cuda:
                self.on_device = True
                return x.data_ptr(), x
            return x.data_ptr(), x
        a = np.ascontiguousarray(x, dtype=dtype)
        self.keep.append(a)
        return a.ctypes.data, a

    def out(self, shape, dtype=np.float32, like=None, device_index=0):
        if like is not None and _is_torch(like) and like.is_cuda:
            tdt = {np.float32: torch.float32, np.int64: torch.int64}[dtype]
            t = torch.empty(tuple(int(s) for s in shape), dtype=tdt, device=like.device)
            self.keep.append(t)
            self.on_device = True
            return t.data_ptr(), t
        a = np.empty(tuple(int(s) for s in shape), dtype=dtype)
        self.keep.append(a)
        return a.ctypes.data, a

    def ctx_for(self, *tensors):
        cuda = [t for t in tensors if _is_torch(t) and t.is_cuda]
        if self.ctx is not None:
            ctx = self.ctx
        else:
            ctx = default_context((cuda[0].device.index or 0) if cuda else 0)
        if cuda:  # order the library's work after torch's on the tensors' stream
            ctx.use_stream(torch.cuda.current_stream(cuda[0].device).cuda_stream)
        return ctx


def _pair_struct(args, in1, in2):
    """in1: [N,]C,H1,W1 (may be a strided view with unit innermost stride), in2: [N,]C,H2,W2."""
    def strides_of(x):
        if _is_torch(x):
            return [int(s) for s in x.stride()], x.data_ptr()
        return [int(s // x.itemsize) for s in x.strides], x.ctypes.data

    def prep(x):
        if _is_torch(x):
            if x.dtype != torch.float32:
                x = x.float()
            if x.dim() == 3:
                x = x.unsqueeze(0)
            # dm_pair reads a stride of 0 as "contiguous default": broadcast / expanded views (stride 0
            # over a dim of size > 1) are materialised instead of being misread as dense
            if x.stride(-1) != 1 or any(s < 0 for s in x.stride()) or \
                    any(s == 0 and n > 1 for s, n in zip(x.stride(), x.shape)):
                x = x.contiguous()
            if x.is_cuda:
                args.on_device = True
        else:
            x = np.asarray(x)
            if x.dtype != np.float32:
                x = x.astype(np.float32)
            if x.ndim == 3:
                x = x[None]
            if x.strides[-1] != x.itemsize or any(s < 0 for s in x.strides) or \
                    any(s == 0 and n > 1 for s, n in zip(x.strides, x.shape)):
                x = np.ascontiguousarray(x)
        args.keep.append(x)
        return x

    a, b = prep(in1), prep(in2)
    if a.ndim != 4 or b.ndim != 4:
        raise DepthMatchError(_lib.DM_ERR_INVALID, "inputs must be [N,]C,H,W")
    if tuple(a.shape[:2]) != tuple(b.shape[:2]):
        raise DepthMatchError(_lib.DM_ERR_INVALID, "in1 and in2 differ in pairs/channels")
    sa, pa = strides_of(a)
    sb, pb = strides_of(b)
    p = dm_pair()
    p.in1, p.in2 = pa, pb
    p.n_pairs, p.channels = int(a.shape[0]), int(a.shape[1])
    p.h1, p.w1, p.h2, p.w2 = int(a.shape[2]), int(a.shape[3]), int(b.shape[2]), int(b.shape[3])
    # a size-1 dim may report any stride (0 included, which dm_pair reads as "default"): give it
    # the stride a dense layout of the inner dims would have
    def dense(st, shape):
        st = list(st)
        for d in (2, 1, 0):
            if shape[d] == 1:
                st[d] = st[d + 1] * int(shape[d + 1])
        return st
    sa, sb = dense(sa, a.shape), dense(sb, b.shape)
    p.in1_stride_n, p.in1_stride_c, p.in1_stride_y = sa[0], sa[1], sa[2]
    p.in2_stride_n, p.in2_stride_c, p.in2_stride_y = sb[0], sb[1], sb[2]
    return p, a, b


# --------------------------------------------------------- low-level wrappers
def match_volume(in1, in2, maxh, maxw, softmax=False, exact=False, ctx=None):
    """nn.SpatialMatching (+ Minus + SoftMax when softmax=True): [N,]H1,W1,maxh,maxw."""
    args = _Args(ctx)
    single = (in1.dim() if _is_torch(in1) else np.ndim(in1)) == 3
    p, a, b = _pair_struct(args, in1, in2)
    c = args.ctx_for(a, b)
    optr, out = args.out((p.n_pairs, p.h1, p.w1, maxh, maxw), np.float32, like=a)
    mode = (DM_VOLUME_NEG_SOFTMAX if softmax else DM_VOLUME_SSD) | (DM_VOLUME_EXACT if exact else 0)
    check(c._lib.dm_match_volume(c.handle, C.byref(p), maxh, maxw, mode, optr))
    return out[0] if single else out


def match_extract_raw_ssd(in1, in2, maxh, maxw, threshold=0.21, ctx=None):
    """extractOutput on the raw SSD volume without the volume: what the ground-truth generators compute
    (radial/radial_opticalflow_groundtruth.lua:105, version2/groundtruth.lua:103).  Returns
    (ret [N,]H1,W1 int64, scores [N,]H1,W1 float32, untouched pixels per pair)."""
    args = _Args(ctx)
    single = (in1.dim() if _is_torch(in1) else np.ndim(in1)) == 3
    p, a, b = _pair_struct(args, in1, in2)
    c = args.ctx_for(a, b)
    rptr, ret = args.out((p.n_pairs, p.h1, p.w1), np.int64, like=a)
    sptr, sc = args.out((p.n_pairs, p.h1, p.w1), np.float32, like=a)
    nun = np.zeros((p.n_pairs,), np.int64)
    check(c._lib.dm_match_extract_raw_ssd(c.handle, C.byref(p), maxh, maxw, float(threshold), rptr, sptr,
                                          nun.ctypes.data))
    if single:
        return ret[0], sc[0], int(nun[0])
    return ret, sc, nun


def match_extract(in1, in2, maxh, maxw, tie_middle=True, exact=False, prob_threshold=0.11,
                  canvas=None, want=("index", "min_ssd", "pmax", "index_thr", "score_thr", "soft_yx"),
                  ctx=None, out=None, async_=False, diff_form=False):
    """Fused prepareInput-less forward + processOutput.  Returns a dict of arrays, each
    [N,]H1,W1 (soft_yx: [N,]2,H1,W1; flow_full: [N,]2,hImg,wImg when canvas=(hImg,wImg)).
    `out` may hold preallocated result buffers by name (e.g. pinned host memory), with the
    batched shapes [N,...]; they are written in place and returned.
    async_=True (host buffers, preallocated `out`): returns once the work is queued; the inputs
    and `out` belong to the library until ctx.synchronize()."""
    args = _Args(ctx)
    single = (in1.dim() if _is_torch(in1) else np.ndim(in1)) == 3
    p, a, b = _pair_struct(args, in1, in2)
    c = args.ctx_for(a, b)
    N, H1, W1 = p.n_pairs, p.h1, p.w1
    o = dm_extract_out()
    res = {}
    shapes = {"index": ((N, H1, W1), np.int64), "min_ssd": ((N, H1, W1), np.float32),
              "pmax": ((N, H1, W1), np.float32), "index_thr": ((N, H1, W1), np.int64),
              "score_thr": ((N, H1, W1), np.float32), "soft_yx": ((N, 2, H1, W1), np.float32),
              "n_untouched": ((N,), np.int64), "conf_marginal": ((N, H1, W1), np.float32)}
    want = list(want)
    if canvas is not None and "flow_full" not in want:
        want.append("flow_full")
    for name in want:
        if name == "flow_full":
            if canvas is None:
                raise DepthMatchError(_lib.DM_ERR_INVALID, "flow_full needs canvas=(hImg, wImg)")
            shape, dt = (N, 2, int(canvas[0]), int(canvas[1])), np.float32
        else:
            shape, dt = shapes[name]
        if out is not None and name in out:
            arr = out[name]
            _check_inplace(arr, dt, shape)
            args.keep.append(arr)
            ptr = _ptr(arr)
        else:
            ptr, arr = args.out(shape, dt, like=a)
        setattr(o, name, ptr)
        res[name] = arr
    flags = (DM_FLAG_TIE_MIDDLE if tie_middle else 0) | (DM_FLAG_EXACT_SSD if exact else 0)
    if diff_form:
        flags |= DM_FLAG_DIFF_SSD
    if async_:
        if out is None or any(n not in out for n in want) or not _caller_owned_host(in1, a, in2, b):
            raise DepthMatchError(_lib.DM_ERR_INVALID, "async_ needs caller-owned `out` buffers for every "
                                  "requested result and inputs that need no conversion copy")
        flags |= DM_FLAG_ASYNC
    himg, wimg = (int(canvas[0]), int(canvas[1])) if canvas is not None else (H1, W1)
    check(c._lib.dm_match_extract(c.handle, C.byref(p), maxh, maxw, flags, float(prob_threshold),
                                  himg, wimg, C.byref(o)))
    if single:
        res = {k: v[0] for k, v in res.items()}
    return res


def _caller_owned_host(in1, a, in2, b):
    """True when the staged views a, b are the caller's own host memory (no conversion copy that
    would be freed while an asynchronous call still reads it)."""
    for x, v in ((in1, a), (in2, b)):
        if not isinstance(x, np.ndarray) or not isinstance(v, np.ndarray) or not np.shares_memory(x, v):
            return False
    return True


# ------------------------------------------------------------------ geometry
class Geometry(dict):
    """The reference's `geometry` table (opticalflow.lua:138-198): attribute access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            return None

    def __setattr__(self, k, v):
        self[k] = v


def round_lua(x):
    """common.lua round(): floor(x + 0.5)."""
    return math.floor(x + 0.5)


# opticalflow_model.lua:12-43
def yx2x(geometry, y, x):
    return (y - 1) * geometry.maxw + x


def x2yx(geometry, x):
    if isinstance(x, (int, float)):
        return math.floor((x - 1) / geometry.maxw) + 1, (x - 1) % geometry.maxw + 1
    xd = (x.double() if _is_torch(x) else np.asarray(x, np.float64)) - 1
    yout = (xd / geometry.maxw).floor() if _is_torch(x) else np.floor(xd / geometry.maxw)
    xout = xd - yout * geometry.maxw
    if _is_torch(x):
        return (yout + 1.5).floor(), (xout + 1.5).floor()
    return np.floor(yout + 1.5), np.floor(xout + 1.5)


def centered2onebased(geometry, y, x):
    return y + math.ceil(geometry.maxh / 2), x + math.ceil(geometry.maxw / 2)


def onebased2centered(geometry, y, x):
    return y - math.ceil(geometry.maxh / 2), x - math.ceil(geometry.maxw / 2)


def getMiddleIndex(geometry):
    if geometry.multiscale:
        return yx2xMulti(geometry, 0, 0)
    y, x = centered2onebased(geometry, 0, 0)
    return yx2x(geometry, y, x)


# opticalflow_model.lua:131-151
def prepareInput(geometry, patch1, patch2):
    if tuple(patch1.shape) != tuple(patch2.shape):
        raise DepthMatchError(_lib.DM_ERR_INVALID, "prepareInput: patches differ in size")
    if geometry.multiscale:
        return [patch1, patch2]
    h, w = patch1.shape[1], patch1.shape[2]
    oy, ox = math.ceil(geometry.maxh / 2) - 1, math.ceil(geometry.maxw / 2) - 1
    return [patch1[:, oy:oy + h - geometry.maxh + 1, ox:ox + w - geometry.maxw + 1], patch2]


# ------------------------------------------------------------------ nn modules
class _Module:
    def forward(self, inp):
        return self.updateOutput(inp)

    __call__ = forward

    def updateGradInput(self, inp, gradOutput):
        raise DepthMatchError(_lib.DM_ERR_UNSUPPORTED,
                              "%s: backward is outside the inference hot path" % type(self).__name__)

    backward = accGradParameters = updateGradInput

    def parameters(self):
        return [], []


class SpatialMatching(_Module):
    """nn.SpatialMatching(maxh, maxw, full_output) (nnx; opticalflow_model.lua:93)."""

    def __init__(self, maxh, maxw, full_output=False, exact=False, ctx=None):
        if full_output:
            raise DepthMatchError(_lib.DM_ERR_UNSUPPORTED, "full_output=true is never used by the reference")
        self.maxh, self.maxw, self.full_output, self.exact, self.ctx = maxh, maxw, False, exact, ctx
        self.output = None

    def updateOutput(self, inp):
        self.output = match_volume(inp[0], inp[1], self.maxh, self.maxw, exact=self.exact, ctx=self.ctx)
        return self.output


class SpatialRadialMatching(_Module):
    """nn.SpatialRadialMatching(hWin) (radial/radial_opticalflow_network.lua:32-34): H1 x W x hWin."""

    def __init__(self, hWin, ctx=None):
        self.hWin, self.ctx = hWin, ctx
        self.output = None

    def updateOutput(self, inp):
        v = match_volume(inp[0], inp[1], self.hWin, 1, exact=True, ctx=self.ctx)
        self.output = v.reshape(tuple(v.shape[:-2]) + (self.hWin,))
        return self.output

    def argmin_flow(self, inp):
        """matcher + `output:min(3)`, idx-1 (radial/test_radial_opticalflow.lua:204-207), fused."""
        args = _Args(self.ctx)
        single = (inp[0].dim() if _is_torch(inp[0]) else np.ndim(inp[0])) == 3
        p, a, b = _pair_struct(args, inp[0], inp[1])
        c = args.ctx_for(a, b)
        fptr, flow = args.out((p.n_pairs, p.h1, p.w1), np.float32, like=a)
        mptr, mn = args.out((p.n_pairs, p.h1, p.w1), np.float32, like=a)
        check(c._lib.dm_radial_match_extract(c.handle, C.byref(p), self.hWin, fptr, mptr))
        return (flow[0], mn[0]) if single else (flow, mn)


class CascadingAddTable(_Module):
    """nn.CascadingAddTable(ratios, trainable, single_beta), forward only (CascadingAddTable.lua)."""

    def __init__(self, ratios, trainable=True, single_beta=False, ctx=None):
        self.ratios, self.trainable, self.ctx = list(ratios), trainable, ctx
        self.output = []

    def updateOutput(self, inp):
        if len(inp) != len(self.ratios):
            raise DepthMatchError(_lib.DM_ERR_INVALID,
                                  "nn.CascadingAddTable: input and ratios must have the same size")
        for t in inp:
            if len(t.shape) != 3:
                raise DepthMatchError(_lib.DM_ERR_INVALID, "nn.CascadingAddTable: input must be a "
                                      "table of 3D-tensors (HxW) x Kh x Kw")
        args = _Args(self.ctx)
        if _is_torch(inp[0]):
            stacked = torch.stack([t.float() for t in inp]).contiguous()
        else:
            stacked = np.stack([np.asarray(t, np.float32) for t in inp])
        iptr, stacked = args.inp(stacked)
        c = args.ctx_for(stacked)
        n, rows, kh, kw = stacked.shape
        optr, out = args.out(stacked.shape, np.float32, like=stacked)
        rat = (C.c_int * n)(*self.ratios)
        check(c._lib.dm_cascade_add(c.handle, iptr, rows, kh, kw, rat, n, optr))
        self.output = [out[i] for i in range(n)]
        return self.output


class OutputExtractor(_Module):
    """nn.OutputExtractor(maxh, maxw): {x, y} soft means (OutputExtractor.lua:21-35)."""

    def __init__(self, maxh, maxw, ctx=None):
        self.maxh, self.maxw, self.ctx = maxh, maxw, ctx
        self.output = None

    def updateOutput(self, inp):
        args = _Args(self.ctx)
        iptr, a = args.inp(inp)
        c = args.ctx_for(a)
        h, w = a.shape[0], a.shape[1]
        yptr, y = args.out((h, w), np.float32, like=a)
        xptr, x = args.out((h, w), np.float32, like=a)
        check(c._lib.dm_soft_mean(c.handle, iptr, h * w, self.maxh, self.maxw, yptr, xptr))
        self.output = [x, y]
        return self.output


# ------------------------------------------------------------------ feature extractor
class _Tables:
    """nn.tables (Torch7 nn): connection tables, rows of 1-based (from, to)."""

    @staticmethod
    def full(nin, nout):
        return np.array([[i + 1, o + 1] for o in range(nout) for i in range(nin)], np.int32)

    @staticmethod
    def random(nin, nout, nto, rng=None):
        """nn.tables.random(nin, nout, nto): every output plane reads `nto` distinct input planes,
        drawn as consecutive chunks of a random permutation (re-drawn when used up).  The
        permutation comes from numpy's generator, not Torch7's Mersenne stream."""
        rng = rng if rng is not None else np.random.default_rng()
        tbl = np.zeros((nout * nto, 2), np.int32)
        nfi = nin // nto
        if nfi < 1:
            raise DepthMatchError(_lib.DM_ERR_INVALID, "nn.tables.random: nto > nin")
        fi, cnt = rng.permutation(nin), 0
        for o in range(nout):
            tbl[o * nto:(o + 1) * nto, 0] = fi[cnt * nto:(cnt + 1) * nto] + 1
            tbl[o * nto:(o + 1) * nto, 1] = o + 1
            cnt += 1
            if cnt == nfi:
                fi, cnt = rng.permutation(nin), 0
        return tbl


class SpatialConvolution(_Module):
    """nn.SpatialConvolution(nInputPlane, nOutputPlane, kW, kH): valid cross-correlation,
    weight [nOut, nIn, kH, kW], bias [nOut]; reset() = uniform(+-1/sqrt(kW*kH*nIn))."""

    def __init__(self, nInputPlane, nOutputPlane, kW, kH, rng=None):
        self.nInputPlane, self.nOutputPlane, self.kW, self.kH = nInputPlane, nOutputPlane, kW, kH
        self.connTable = None
        self.reset(rng)

    def reset(self, rng=None):
        rng = rng if rng is not None else np.random.default_rng()
        stdv = 1.0 / math.sqrt(self.kW * self.kH * self.nInputPlane)
        self.weight = rng.uniform(-stdv, stdv, (self.nOutputPlane, self.nInputPlane, self.kH, self.kW)).astype(np.float32)
        self.bias = rng.uniform(-stdv, stdv, self.nOutputPlane).astype(np.float32)

    def updateOutput(self, inp):
        return Filter([self]).forward(inp)


class SpatialConvolutionMap(_Module):
    """nn.SpatialConvolutionMap(connTable, kW, kH): weight [nConn, kH, kW], bias [nOut]."""

    def __init__(self, connTable, kW, kH, rng=None):
        self.connTable = np.ascontiguousarray(connTable, np.int32)
        self.kW, self.kH = kW, kH
        self.nInputPlane = int(self.connTable[:, 0].max())
        self.nOutputPlane = int(self.connTable[:, 1].max())
        self.reset(rng)

    def reset(self, rng=None):
        rng = rng if rng is not None else np.random.default_rng()
        n = self.connTable.shape[0]
        self.weight = np.empty((n, self.kH, self.kW), np.float32)
        self.bias = np.empty(self.nOutputPlane, np.float32)
        for o in range(self.nOutputPlane):
            rows = np.nonzero(self.connTable[:, 1] == o + 1)[0]
            stdv = 1.0 / math.sqrt(self.kW * self.kH * max(len(rows), 1))
            self.weight[rows] = rng.uniform(-stdv, stdv, (len(rows), self.kH, self.kW))
            self.bias[o] = rng.uniform(-stdv, stdv)

    def updateOutput(self, inp):
        return Filter([self]).forward(inp)


class Tanh(_Module):
    def updateOutput(self, inp):
        raise DepthMatchError(_lib.DM_ERR_UNSUPPORTED, "nn.Tanh runs fused behind a convolution (Filter)")


class Filter(_Module):
    """The nn.Sequential getFilter returns (opticalflow_model.lua:45-79): convolution layers with
    nn.Tanh between them.  forward(input): [C,H,W] or a batch [N,C,H,W] (numpy or torch CUDA);
    pads = (l, r, t, b) is an nn.SpatialZeroPadding in front (multiscale prefilter).  The weights
    are packed on the device at the first forward; call reset_weights() after editing them."""

    def __init__(self, modules, ctx=None):
        self.modules, self.ctx = list(modules), ctx
        self._handle, self._handle_ctx, self._handle_cin = None, None, None
        self.output = None

    def add(self, m):
        self.modules.append(m)
        self.reset_weights()
        return self

    def getWeights(self):
        convs = [m for m in self.modules if hasattr(m, "weight")]
        return {"layer%d" % (i + 1): m.weight for i, m in enumerate(convs)}

    def reset_weights(self):
        if self._handle is not None:
            self._handle_ctx._lib.dm_filter_destroy(self._handle)
        self._handle, self._handle_ctx = None, None

    def __del__(self):
        try:  # interpreter shutdown may have torn the library or the context down already
            self.reset_weights()
        except Exception:
            pass

    def _layers(self):
        out = []
        for m in self.modules:
            if isinstance(m, Tanh):
                if not out:
                    raise DepthMatchError(_lib.DM_ERR_UNSUPPORTED, "a Filter starts with a convolution")
                out[-1][1] = 1
            else:
                out.append([m, 0])
        return out

    def _build(self, c, cin):
        if self._handle is not None and self._handle_ctx is c and self._handle_cin == cin:
            return self._handle
        self.reset_weights()
        layers = self._layers()
        arr = (_lib.dm_conv_layer * len(layers))()
        keep = []
        planes = cin
        for i, (m, tanh) in enumerate(layers):
            w = np.ascontiguousarray(m.weight, np.float32)
            b = np.ascontiguousarray(m.bias, np.float32)
            keep += [w, b]
            # a connection table may leave the last input planes unused (nn.tables.random):
            # like Torch7, accept an input with more planes than the table mentions
            n_in = planes if m.connTable is not None and planes >= m.nInputPlane else m.nInputPlane
            arr[i].n_in, arr[i].n_out, arr[i].kh, arr[i].kw = n_in, m.nOutputPlane, m.kH, m.kW
            planes = m.nOutputPlane
            arr[i].tanh_after = tanh
            arr[i].weight, arr[i].bias = w.ctypes.data, b.ctypes.data
            if m.connTable is not None:
                t = np.ascontiguousarray(m.connTable, np.int32)
                keep.append(t)
                arr[i].n_conn, arr[i].conn = t.shape[0], t.ctypes.data
                if w.shape != (t.shape[0], m.kH, m.kW):
                    raise DepthMatchError(_lib.DM_ERR_INVALID, "SpatialConvolutionMap: weight shape")
            elif w.shape != (m.nOutputPlane, m.nInputPlane, m.kH, m.kW):
                raise DepthMatchError(_lib.DM_ERR_INVALID, "SpatialConvolution: weight shape")
        h = C.c_void_p()
        check(c._lib.dm_filter_create(c.handle, arr, len(layers), C.byref(h)))
        self._handle, self._handle_ctx, self._handle_cin = h, c, cin
        return h

    def output_size(self, h, w, pads=(0, 0, 0, 0)):
        for m in self.modules:
            if hasattr(m, "weight"):
                h, w = h - m.kH + 1, w - m.kW + 1
        return h + pads[2] + pads[3], w + pads[0] + pads[1]

    def updateOutput(self, inp, pads=(0, 0, 0, 0)):
        args = _Args(self.ctx)
        iptr, a = args.inp(inp)
        c = args.ctx_for(a)
        single = len(a.shape) == 3
        shp = (1,) + tuple(a.shape) if single else tuple(a.shape)
        n, cin, h, w = shp
        convs = [m for m in self.modules if hasattr(m, "weight")]
        if cin != convs[0].nInputPlane and not (convs[0].connTable is not None and cin > convs[0].nInputPlane):
            raise DepthMatchError(_lib.DM_ERR_INVALID, "Filter: %d input planes, the first layer takes %d"
                                  % (cin, convs[0].nInputPlane))
        fh = self._build(c, cin)
        co, ho, wo = C.c_int(), C.c_int(), C.c_int()
        check(c._lib.dm_filter_output_size(fh, h, w, *[int(v) for v in pads], C.byref(co), C.byref(ho),
                                           C.byref(wo)))
        optr, out = args.out((n, co.value, ho.value, wo.value), np.float32, like=a)
        check(c._lib.dm_filter_forward(c.handle, fh, iptr, n, h, w, *[int(v) for v in pads], optr))
        self.output = out[0] if single else out
        return self.output

    def forward(self, inp, pads=(0, 0, 0, 0)):
        return self.updateOutput(inp, pads)

    __call__ = forward


# opticalflow_model.lua:45-79 (layers[i] = {nIn, kW, kH, nOut}, tanh between layers)
def getFilter(geometry, rng=None, ctx=None):
    if geometry.L2Pooling:
        raise DepthMatchError(_lib.DM_ERR_UNSUPPORTED, "getFilter: L2Pooling (the reference asserts too)")
    rng = rng if rng is not None else np.random.default_rng()
    mods, layers = [], geometry.layers
    for i, l in enumerate(layers):
        if i == 0 or layers[i - 1][3] == l[0]:
            mods.append(SpatialConvolution(l[0], l[3], l[1], l[2], rng))
        else:
            mods.append(SpatialConvolutionMap(_Tables.random(layers[i - 1][3], l[3], l[0], rng), l[1], l[2], rng))
        if i != len(layers) - 1:
            mods.append(Tanh())
    return Filter(mods, ctx=ctx)


# radial/radial_opticalflow_network.lua:6-31 (layers: 'tanh' or {nIn, kH, kW, nOut})
def getRadialFilter(networkp, rng=None, ctx=None):
    rng = rng if rng is not None else np.random.default_rng()
    mods, last = [], None
    for l in networkp["layers"]:
        if isinstance(l, str):
            if l != "tanh":
                raise DepthMatchError(_lib.DM_ERR_INVALID, "Unknown layer %r" % (l,))
            mods.append(Tanh())
        elif last is None or l[0] == last:
            mods.append(SpatialConvolution(l[0], l[3], l[2], l[1], rng))
            last = l[3]
        else:
            mods.append(SpatialConvolutionMap(_Tables.random(last, l[3], l[0], rng), l[2], l[1], rng))
            last = l[3]
    return Filter(mods, ctx=ctx)


# opticalflow_model_multiscale.lua:134-173
def getMultiscalePrefilter(geometry, filter, ctx=None):
    """Returns prefilter(img [C,H,W]) -> one feature map per ratio: r x r average
    (nn.SpatialDownSampling), zero padding of the patch footprint, the (shared) filter."""
    wPad, hPad = geometry.wPatch2 - 1, geometry.hPatch2 - 1
    pads = (wPad // 2, wPad - wPad // 2, hPad // 2, hPad - hPad // 2)
    filters = [filter if (geometry.share_filters or i == 0) else Filter(filter.modules, ctx=ctx)
               for i in range(len(geometry.ratios))]

    def prefilter(img):
        outs = []
        for r, f in zip(geometry.ratios, filters):
            outs.append(f.forward(downsample(img, r, ctx=ctx) if r != 1 else img, pads))
        return outs

    prefilter.getWeights = filter.getWeights
    return prefilter


def multiscaleInputs(geometry, filter, img1, img2, ctx=None):
    """The per-scale {f1, f2} pairs getModelMultiscale(prefiltered) takes, from two raw frames:
    every scale is the r x r average of the frame, zero-padded by the filter footprint
    (getMultiscalePrefilter) -- frame 2 by the search window on top, so that f2 is
    (maxh-1, maxw-1) larger than f1 and every displacement of the window is a valid read --
    and run through the shared filter."""
    wPad, hPad = geometry.wPatch2 - 1, geometry.hPatch2 - 1
    p1 = (wPad // 2, wPad - wPad // 2, hPad // 2, hPad - hPad // 2)
    wl, wt = (geometry.maxw - 1) // 2, (geometry.maxh - 1) // 2
    p2 = (p1[0] + wl, p1[1] + geometry.maxw - 1 - wl, p1[2] + wt, p1[3] + geometry.maxh - 1 - wt)
    out = []
    for r in geometry.ratios:
        a = downsample(img1, r, ctx=ctx) if r != 1 else img1
        b = downsample(img2, r, ctx=ctx) if r != 1 else img2
        out.append((filter.forward(a, p1), filter.forward(b, p2)))
    return out


def downsample(img, r, ctx=None):
    """nn.SpatialDownSampling(r, r): r x r average (opticalflow_model_multiscale.lua:145)."""
    args = _Args(ctx)
    iptr, a = args.inp(img)
    c = args.ctx_for(a)
    ch, h, w = a.shape
    optr, out = args.out((ch, h // r, w // r), np.float32, like=a)
    check(c._lib.dm_downsample_avg(c.handle, iptr, ch, h, w, r, optr))
    return out


class _NN:
    SpatialMatching = SpatialMatching
    SpatialRadialMatching = SpatialRadialMatching
    CascadingAddTable = CascadingAddTable
    OutputExtractor = OutputExtractor
    SpatialConvolution = SpatialConvolution
    SpatialConvolutionMap = SpatialConvolutionMap
    Tanh = Tanh
    tables = _Tables


nn = _NN()


class _ExtractOutput:
    """`require 'extractoutput'` (extract_output.cpp:357-366): in-place on ret/scores."""

    @staticmethod
    def extractOutput(inp, scores, threshold, ret, ctx=None):
        args = _Args(ctx)
        iptr, a = args.inp(inp)
        c = args.ctx_for(a)
        h, w, n = a.shape
        _check_inplace(scores, np.float32, (h, w))
        _check_inplace(ret, np.int64, (h, w))
        nun = C.c_int64(0)
        check(c._lib.dm_extract_output(c.handle, iptr, h, w, n, float(threshold), _ptr(ret),
                                       _ptr(scores), C.byref(nun)))
        return int(nun.value)

    @staticmethod
    def extractOutputMarginalized(inp, threshold, threshold_acc, ret, retgd, ctx=None):
        args = _Args(ctx)
        iptr, a = args.inp(inp)
        c = args.ctx_for(a)
        h, w, n = a.shape
        _check_inplace(ret, np.int64, (h, w))
        _check_inplace(retgd, np.int64, (h, w))
        check(c._lib.dm_extract_output_marginalized(c.handle, iptr, h, w, n, float(threshold),
                                                    float(threshold_acc), _ptr(ret), _ptr(retgd)))


extractoutput = _ExtractOutput()


def _ptr(x):
    return x.data_ptr() if _is_torch(x) else x.ctypes.data


def _check_inplace(x, dtype, shape):
    if _is_torch(x):
        ok = x.is_contiguous() and tuple(x.shape) == tuple(shape) and \
            x.dtype == {np.float32: torch.float32, np.int64: torch.int64}[dtype]
    else:
        ok = isinstance(x, np.ndarray) and x.flags.c_contiguous and x.dtype == dtype and \
            tuple(x.shape) == tuple(shape)
    if not ok:
        raise DepthMatchError(_lib.DM_ERR_INVALID, "output tensor must be contiguous %s of shape %s"
                              % (np.dtype(dtype).name, tuple(shape)))


# ------------------------------------------------------------------ models
class _MatchModel(_Module):
    """getModel(geometry, full_image, prefiltered=true) (opticalflow_model.lua:81-129):
    SpatialMatching -> Minus -> SoftMax [-> OutputExtractor].  forward() returns what the
    reference's model:forward returns: the H1 x W1 x (maxh*maxw) probability volume."""

    def __init__(self, geometry, ctx=None):
        self.geometry, self.ctx = geometry, ctx
        self.output = None

    def updateOutput(self, inp):
        g = self.geometry
        v = match_volume(inp[0], inp[1], g.maxh, g.maxw, softmax=True, ctx=self.ctx)
        v = v.reshape(tuple(v.shape[:-2]) + (g.maxh * g.maxw,))
        if g.output_extraction_method == "mean":
            self.output = OutputExtractor(g.maxh, g.maxw, ctx=self.ctx).forward(v)
        else:
            self.output = v
        return self.output


class FusedOutput(dict):
    """What DenseMatch.forward returns: per-pixel results, no volume."""


class DenseMatch(_Module):
    """nn.DenseMatch: the fused replacement for getModel's graph + processOutput's reductions
    (depth_estimation_api.lua:164-168 become one call)."""

    def __init__(self, geometry, exact=False, ctx=None):
        self.geometry, self.exact, self.ctx = geometry, exact, ctx
        self.output = None

    def updateOutput(self, inp):
        g = self.geometry
        canvas = (g.hImg, g.wImg) if g.hImg and g.wImg else None
        self.output = FusedOutput(match_extract(inp[0], inp[1], g.maxh, g.maxw, tie_middle=True,
                                                exact=self.exact, canvas=canvas, ctx=self.ctx,
                                                want=("index", "pmax", "index_thr", "score_thr",
                                                      "soft_yx", "conf_marginal")))
        return self.output


class _FilteredModel(_Module):
    """getModel(geometry, full_image, prefiltered=false) (opticalflow_model.lua:81-92): the
    shared-weight filter on both patches (nn.ParallelTable of a filter and its clone), then the
    matcher.  forward({patch1, patch2}) as prepareInput returns them."""

    def __init__(self, geometry, matcher, filter, ctx=None):
        self.geometry, self.matcher, self.filter, self.ctx = geometry, matcher, filter, ctx
        self.modules = [[filter, filter], matcher]
        self.output = None

    def getWeights(self):
        return self.filter.getWeights()

    def updateOutput(self, inp):
        p1, p2 = inp
        if tuple(p1.shape) == tuple(p2.shape):  # one launch per layer for both frames
            both = torch.stack([p1, p2]) if _is_torch(p1) else np.stack([p1, p2])
            f = self.filter.forward(both)
            f1, f2 = f[0], f[1]
        else:
            c = (lambda x: x.contiguous()) if _is_torch(p1) else np.ascontiguousarray
            f1, f2 = self.filter.forward(c(p1)), self.filter.forward(c(p2))
        self.output = self.matcher.forward([f1, f2])
        return self.output


def getModel(geometry, full_image=True, prefiltered=False, fused=False, filter=None, rng=None, ctx=None):
    if geometry.multiscale:
        return getModelMultiscale(geometry, full_image, prefiltered, filter=filter, rng=rng, ctx=ctx)
    matcher = DenseMatch(geometry, ctx=ctx) if fused else _MatchModel(geometry, ctx=ctx)
    if prefiltered:
        return matcher
    return _FilteredModel(geometry, matcher, filter or getFilter(geometry, rng, ctx), ctx=ctx)


# opticalflow_model.lua:153-169
def getOutputConfidences(geometry, inp, threshold=None, ctx=None):
    args = _Args(ctx)
    iptr, a = args.inp(inp)
    c = args.ctx_for(a)
    h, w, K = a.shape
    if threshold is None:
        iptr2, idx = args.out((h, w), np.int64, like=a)
        check(c._lib.dm_argmax_tie(c.handle, iptr, h * w, K, int(getMiddleIndex(geometry)), 0, iptr2,
                                   None))
        ones = torch.ones((h, w), device=a.device) if _is_torch(a) else np.ones((h, w), np.float32)
        return idx, ones
    # the reference allocates uninitialised tensors here (:163-164); we start from zeros
    imaxs = torch.zeros((h, w), dtype=torch.int64, device=a.device) if _is_torch(a) else np.zeros((h, w), np.int64)
    scores = torch.zeros((h, w), device=a.device) if _is_torch(a) else np.zeros((h, w), np.float32)
    extractoutput.extractOutput(a, scores, 0.11, imaxs, ctx=c)
    return imaxs, scores > threshold


# opticalflow_model.lua:171-199
def getOutputConfidences2(geometry, inp, ctx=None):
    x, y = OutputExtractor(geometry.maxh, geometry.maxw, ctx=ctx).forward(inp)
    args = _Args(ctx)
    iptr, a = args.inp(inp)
    c = args.ctx_for(a)
    h, w, K = a.shape
    mptr, pm = args.out((h, w, geometry.maxh), np.float32, like=a)
    check(c._lib.dm_marginal_x(c.handle, iptr, h * w, geometry.maxh, geometry.maxw, mptr))
    imaxs = torch.zeros((h, w), dtype=torch.int64, device=a.device) if _is_torch(a) else np.zeros((h, w), np.int64)
    scores = torch.zeros((h, w), device=a.device) if _is_torch(a) else np.zeros((h, w), np.float32)
    extractoutput.extractOutput(pm, scores, 0.11, imaxs, ctx=c)
    return y, x, scores > 0


# opticalflow_model.lua:201-252
def processOutput(geometry, output, process_full=None, threshold=None, ctx=None):
    ret = {}
    if isinstance(output, MultiscaleOutput):
        ret["index"], ret["y"], ret["x"] = output["index"], output["flow_y"], output["flow_x"]
        ret["confidences"] = _ones_like(ret["index"])
    elif isinstance(output, FusedOutput):
        yoff, xoff = centered2onebased(geometry, 0, 0)
        if geometry.output_extraction_method == "mean":
            ret["y"], ret["x"] = output["soft_yx"][..., 0, :, :] - yoff, output["soft_yx"][..., 1, :, :] - xoff
            fy = _floor(output["soft_yx"][..., 0, :, :] + 0.5)
            fx = _floor(output["soft_yx"][..., 1, :, :] + 0.5)
            ret["index"] = _to_long(yx2x(geometry, fy, fx))
            if "conf_marginal" in output:     # getOutputConfidences2's marginal test, fused
                ret["confidences"] = output["conf_marginal"] > 0
            else:
                ret["confidences"] = _ones_like(ret["index"])
        else:
            if threshold is None:
                ret["index"] = output["index"]
                ret["confidences"] = _ones_like(ret["index"])
            else:
                ret["index"] = output["index_thr"]
                ret["confidences"] = output["score_thr"] > threshold
            ry, rx = x2yx(geometry, ret["index"])
            ret["y"], ret["x"] = ry - yoff, rx - xoff
    elif geometry.output_extraction_method == "max" or geometry.output_extraction_method is None:
        ret["index"], ret["confidences"] = getOutputConfidences(geometry, output, threshold, ctx=ctx)
        if geometry.multiscale:
            ret["y"], ret["x"] = x2yxMulti(geometry, ret["index"], ctx=ctx)
        else:
            ry, rx = x2yx(geometry, ret["index"])
            yoff, xoff = centered2onebased(geometry, 0, 0)
            ret["y"], ret["x"] = ry - yoff, rx - xoff
    else:
        ret["y"], ret["x"], ret["confidences"] = getOutputConfidences2(geometry, output, ctx=ctx)
        ret["index"] = _to_long(yx2x(geometry, _floor(ret["y"] + 0.5), _floor(ret["x"] + 0.5)))
        yoff, xoff = centered2onebased(geometry, 0, 0)
        ret["y"], ret["x"] = ret["y"] - yoff, ret["x"] - xoff
    if process_full is None:
        process_full = True
    if process_full:
        h, w = ret["y"].shape[-2], ret["y"].shape[-1]
        hoff = (geometry.hImg - h) // 2
        woff = (geometry.wImg - w) // 2
        if _is_torch(ret["y"]):
            full = torch.zeros((2, geometry.hImg, geometry.wImg), device=ret["y"].device)
            fc = torch.zeros((geometry.hImg, geometry.wImg), device=ret["y"].device)
        else:
            full = np.zeros((2, geometry.hImg, geometry.wImg), np.float32)
            fc = np.zeros((geometry.hImg, geometry.wImg), np.float32)
        full[0, hoff:hoff + h, woff:woff + w] = ret["y"]
        full[1, hoff:hoff + h, woff:woff + w] = ret["x"]
        fc[hoff:hoff + h, woff:woff + w] = ret["confidences"]
        ret["full"], ret["full_confidences"] = full, fc
    return ret


def _ones_like(x):
    return torch.ones(x.shape, device=x.device) if _is_torch(x) else np.ones(x.shape, np.float32)


def _floor(x):
    return x.floor() if _is_torch(x) else np.floor(x)


def _to_long(x):
    return x.long() if _is_torch(x) else np.asarray(x, np.int64)


# ------------------------------------------------------------------ multiscale
def _ring_d(geometry, i):
    r = geometry.ratios
    return round_lua(geometry.maxw * (r[i] - r[i - 1]) / (2 * r[i]))


def multiscaleLength(geometry):
    L = geometry.maxh * geometry.maxw
    for i in range(1, len(geometry.ratios)):
        d = _ring_d(geometry, i)
        L += 2 * d * geometry.maxw + 2 * (geometry.maxh - 2 * d) * d
    return L


# opticalflow_model_multiscale.lua:10-52
def yx2xMulti(geometry, y, x):
    x, y = round_lua(x), round_lua(y)
    maxh, maxw, ratios = geometry.maxh, geometry.maxw, geometry.ratios

    def is_in(size, v):
        return -math.ceil(size / 2) + 1 <= v <= math.floor(size / 2)

    for i, r in enumerate(ratios):
        if is_in(maxw * r, x) and is_in(maxh * r, y):
            tx = math.ceil(x / r) + math.ceil(maxw / 2)
            ty = math.ceil(y / r) + math.ceil(maxh / 2)
            break
    else:
        raise AssertionError("yx2xMulti: (%d,%d) outside every scale" % (y, x))
    if i == 0:
        return (ty - 1) * maxw + tx
    d = _ring_d(geometry, i)
    if ty <= d:
        it = (ty - 1) * maxw + tx
    elif ty > maxh - d:
        it = d * maxw + 2 * (maxh - 2 * d) * d + (ty - (maxh - d) - 1) * maxw + tx
    elif tx <= d:
        it = d * maxw + (ty - d - 1) * d + tx
    elif tx > maxw - d:
        it = d * maxw + (maxh - 2 * d) * d + (ty - d - 1) * d + tx - (maxw - d)
    else:
        raise AssertionError("yx2xMulti: (%d,%d) falls in the removed middle" % (y, x))
    return maxw * maxh + (i - 1) * (2 * d * maxw + 2 * (maxh - 2 * d) * d) + it


# opticalflow_model_multiscale.lua:83-132
def x2yxMultiNumber(geometry, x):
    maxh, maxw, ratios = geometry.maxh, geometry.maxw, geometry.ratios
    cy, cx = math.ceil(maxh / 2), math.ceil(maxw / 2)
    if x <= maxh * maxw:
        return (x - 1) // maxw + 1 - cy, (x - 1) % maxw + 1 - cx
    x -= maxh * maxw
    for i in range(1, len(ratios)):
        d = _ring_d(geometry, i)
        ln = 2 * d * maxw + 2 * (maxh - 2 * d) * d
        side = (maxh - 2 * d) * d
        if x > ln:
            x -= ln
            continue
        if x <= d * maxw:
            ty, tx = (x - 1) // maxw + 1, (x - 1) % maxw + 1
        elif x - d * maxw <= side:
            x -= d * maxw
            ty, tx = (x - 1) // d + 1 + d, (x - 1) % d + 1
        elif x - d * maxw - side <= side:
            x -= d * maxw + side
            ty, tx = (x - 1) // d + 1 + d, (x - 1) % d + 1 + maxw - d
        else:
            x -= d * maxw + 2 * side
            assert x <= d * maxw
            ty, tx = (x - 1) // maxw + 1 + maxh - d, (x - 1) % maxw + 1
        return (ty - cy) * ratios[i], (tx - cx) * ratios[i]
    raise AssertionError("x2yxMultiNumber: index beyond the last ring")


def x2yxMulti2(geometry, x, bug_compat=False, ctx=None):
    """The vectorised decode (opticalflow_model_multiscale.lua:72-81): returns (rety, retx)."""
    args = _Args(ctx)
    xptr, a = args.inp(x, np.int64)
    c = args.ctx_for(a)
    h, w = (a.shape[0], a.shape[1]) if a.ndim == 2 else (1, int(np.prod(a.shape)))
    if bug_compat:  # entries the C falls through on keep their previous content: start from 0
        rety = torch.zeros_like(a) if _is_torch(a) else np.zeros_like(a)
        retx = torch.zeros_like(a) if _is_torch(a) else np.zeros_like(a)
        yptr, xptr2 = _ptr(rety), _ptr(retx)
    else:
        yptr, rety = args.out(a.shape, np.int64, like=a)
        xptr2, retx = args.out(a.shape, np.int64, like=a)
    rat = (C.c_int * len(geometry.ratios))(*geometry.ratios)
    check(c._lib.dm_x2yx_multi(c.handle, xptr, h, w, geometry.maxh, geometry.maxw, rat,
                               len(geometry.ratios), 1 if bug_compat else 0, yptr, xptr2))
    return rety, retx


def x2yxMulti(geometry, x, ctx=None):
    if isinstance(x, (int, float)):
        return x2yxMultiNumber(geometry, int(x))
    return x2yxMulti2(geometry, x, ctx=ctx)


class MultiscaleOutput(dict):
    pass


class _MultiscaleModel(_Module):
    """getModelMultiscale(geometry, full_image, prefiltered=true)
    (opticalflow_model_multiscale.lua:175-373) + the argmax/decode of processOutput, fused.
    forward(input): input[i] = {f1_i, f2_i}, the per-scale prefiltered maps
    (f1_i: C x H/r x W/r after the window crop, f2_i: C x (H/r+maxh-1) x (W/r+maxw-1))."""

    def __init__(self, geometry, ctx=None):
        self.geometry, self.ctx = geometry, ctx
        self.output = None

    def updateOutput(self, inp):
        g = self.geometry
        n = len(g.ratios)
        if len(inp) != n:
            raise DepthMatchError(_lib.DM_ERR_INVALID, "one {f1,f2} pair per ratio is required")
        args = _Args(self.ctx)
        p1 = (C.c_void_p * n)()
        p2 = (C.c_void_p * n)()
        first = None
        for i, (f1, f2) in enumerate(inp):
            a1, t1 = args.inp(f1)
            a2, t2 = args.inp(f2)
            p1[i], p2[i] = a1, a2
            first = first if first is not None else t1
        c = args.ctx_for(first)
        ch = first.shape[0]
        h, w = first.shape[1] * g.ratios[0], first.shape[2] * g.ratios[0]
        iptr, idx = args.out((h, w), np.int64, like=first)
        yptr, fy = args.out((h, w), np.int64, like=first)
        xptr, fx = args.out((h, w), np.int64, like=first)
        rat = (C.c_int * n)(*g.ratios)
        check(c._lib.dm_multiscale_extract(c.handle, p1, p2, ch, h, w, g.maxh, g.maxw, rat, n, iptr,
                                           yptr, xptr))
        self.output = MultiscaleOutput(index=idx, flow_y=fy, flow_x=fx)
        return self.output


class _MultiscaleFromFrames(_Module):
    """getModelMultiscale(geometry, full_image, prefiltered=false): the per-scale prefilter
    (average, zero padding, shared filter -- multiscaleInputs) in front of the prefiltered model.
    forward({frame1, frame2}), both [C,H,W] with H, W multiples of the largest ratio.  The border
    handling of the reference's nnx SpatialPyramid (out of tree) is replaced by zero padding of
    the frames, see multiscaleInputs."""

    def __init__(self, geometry, filter, ctx=None):
        self.geometry, self.filter, self.ctx = geometry, filter, ctx
        self.inner = _MultiscaleModel(geometry, ctx=ctx)
        self.output = None

    def getWeights(self):
        return self.filter.getWeights()

    def updateOutput(self, inp):
        self.output = self.inner.forward(multiscaleInputs(self.geometry, self.filter, inp[0], inp[1], ctx=self.ctx))
        return self.output


def getModelMultiscale(geometry, full_image=True, prefiltered=False, filter=None, rng=None, ctx=None):
    assert geometry.output_extraction_method in (None, "max")
    assert geometry.ratios[0] == 1
    if prefiltered:
        return _MultiscaleModel(geometry, ctx=ctx)
    return _MultiscaleFromFrames(geometry, filter or getFilter(geometry, rng, ctx), ctx=ctx)


# ------------------------------------------------------------------ radial
# radial/radial_opticalflow_polar.lua:4-10
def getRMax(h, w, e2):
    ex, ey = float(e2[0]), float(e2[1])
    return math.floor(math.sqrt(max(max(ex * ex + ey * ey, (w - ex) ** 2 + ey * ey),
                                    max(ex * ex + (h - ey) ** 2, (w - ex) ** 2 + (h - ey) ** 2))))


# radial/cartesian2polar.lua:4-49
def getC2PMask(wsrc, hsrc, wdst, hdst, xcenter=None, ycenter=None, lpadding=0, rpadding=0, rmax=None,
               alpha=1.0, ctx=None):
    rmax = rmax if rmax is not None else min(hsrc // 2, wsrc // 2) - 1
    xcenter = wsrc / 2 if xcenter is None else xcenter
    ycenter = hsrc / 2 if ycenter is None else ycenter
    c = ctx or default_context()
    mask = np.empty((2, hdst, wdst + lpadding + rpadding), np.float32)
    check(c._lib.dm_c2p_mask(c.handle, wdst, hdst, float(xcenter), float(ycenter), lpadding, rpadding,
                             float(rmax), float(alpha), mask.ctypes.data))
    return mask


# radial/cartesian2polar.lua:51-89
def getP2CMask(wsrc, hsrc, wdst, hdst, xcenter=None, ycenter=None, rmax=None, alpha=1.0, ctx=None):
    wdst, hdst = int(wdst), int(hdst)  # torch.FloatTensor(2, hdst, wdst) truncates
    rmax = rmax if rmax is not None else min(hdst // 2, wdst // 2) - 1
    xcenter = wdst / 2 if xcenter is None else xcenter
    ycenter = hdst / 2 if ycenter is None else ycenter
    c = ctx or default_context()
    mask = np.empty((2, hdst, wdst), np.float32)
    check(c._lib.dm_p2c_mask(c.handle, int(wsrc), int(hsrc), wdst, hdst, float(xcenter),
                             float(ycenter), float(rmax), float(alpha), mask.ctypes.data))
    return mask


# radial/cartesian2polar.lua:91-93
def cartesian2polar(img, mask=None, ctx=None, **analytic):
    """With `mask`: image.warp(img, mask, 'bilinear', false).  Without: the analytic remap,
    keyword arguments as getC2PMask (wdst, hdst, xcenter, ycenter, lpadding, rpadding, rmax, alpha)."""
    args = _Args(ctx)
    squeeze = (img.dim() if _is_torch(img) else np.ndim(img)) == 2
    if squeeze:
        img = img[None]
    sptr, s = args.inp(img)
    c = args.ctx_for(s)
    ch, hs, ws = s.shape
    if mask is not None:
        mptr, m = args.inp(mask)
        hd, wd = m.shape[1], m.shape[2]
        optr, out = args.out((ch, hd, wd), np.float32, like=s)
        check(c._lib.dm_warp_bilinear(c.handle, sptr, ch, hs, ws, mptr, hd, wd, optr))
    else:
        wdst, hdst = analytic["wdst"], analytic["hdst"]
        lp, rp = analytic.get("lpadding", 0), analytic.get("rpadding", 0)
        optr, out = args.out((ch, hdst, wdst + lp + rp), np.float32, like=s)
        check(c._lib.dm_polar_remap(c.handle, sptr, ch, hs, ws, float(analytic["xcenter"]),
                                    float(analytic["ycenter"]), float(analytic["rmax"]),
                                    float(analytic.get("alpha", 1.0)), lp, rp, optr, hdst, wdst))
    return out[0] if squeeze else out


def polar2cartesian(polar, wdst, hdst, xcenter, ycenter, rmax, alpha=1.0, ctx=None):
    """cartesian2polar(polar, getP2CMask(wsrc, hsrc, wdst, hdst, ...)) without the LUT."""
    args = _Args(ctx)
    squeeze = (polar.dim() if _is_torch(polar) else np.ndim(polar)) == 2
    if squeeze:
        polar = polar[None]
    sptr, s = args.inp(polar)
    c = args.ctx_for(s)
    ch, hs, ws = s.shape
    optr, out = args.out((ch, int(hdst), int(wdst)), np.float32, like=s)
    check(c._lib.dm_polar_unmap(c.handle, sptr, ch, hs, ws, float(xcenter), float(ycenter),
                                float(rmax), float(alpha), optr, int(hdst), int(wdst)))
    return out[0] if squeeze else out


# radial/radial_opticalflow_polar.lua:12-30
def getKOutput(networkp):
    hPolar = networkp["hInput"] - math.floor((networkp["hKernel"] - 1) / 2) - networkp["hWin"] + 1
    return hPolar / networkp["hInput"]


def getP2CMaskOF(networkp, e2, alpha_polar=None, ctx=None):
    wPolar = networkp["wInput"]
    hPolar = networkp["hInput"] - networkp["hKernel"] - networkp["hWin"] + 2
    kOutput = hPolar / networkp["hInput"]
    wOutput, hOutput = networkp["wImg"] * kOutput, networkp["hImg"] * kOutput
    new_e2 = (e2[0] * kOutput, e2[1] * kOutput)
    newRMax = getRMax(networkp["hImg"], networkp["wImg"], e2) * kOutput
    return getP2CMask(wPolar, hPolar, wOutput, hOutput, new_e2[0], new_e2[1], newRMax,
                      1.0 if alpha_polar is None else alpha_polar, ctx=ctx)


# radial/radial_opticalflow_display.lua:6-58
def flow2depth(networkp, flow, center=None, kinfty=0.65, ctx=None):
    args = _Args(ctx)
    fptr, f = args.inp(flow)
    c = args.ctx_for(f)
    h, w = f.shape
    if center is None:
        center = (w / 2, h / 2)
    infty = getRMax(networkp["hImg"], networkp["wImg"], center) * kinfty
    dptr, depth = args.out((h, w), np.float32, like=f)
    cptr, confs = args.out((h, w), np.float32, like=f)
    check(c._lib.dm_flow2depth(c.handle, fptr, h, w, float(center[0]), float(center[1]), float(infty),
                               dptr, cptr))
    return depth / infty, confs


# ------------------------------------------------------------------ next rows
# opticalflow_model.lua:323-472
def postProcessImage(inp, mask, winsize, method, ctx=None):
    args = _Args(ctx)
    iptr, a = args.inp(inp)
    mptr, m = args.inp(mask)
    c = args.ctx_for(a, m)
    _, h, w = a.shape
    optr, out = args.out((2, h, w), np.float32, like=a)
    check(c._lib.dm_post_process_image(c.handle, iptr, mptr, h, w, int(winsize),
                                       1 if method == "max" else 0, optr))
    return out


# depth_estimation_api.lua:76-132 (in place, returns the mask like the reference)
def enlargeMask(mask, ix, iy, ctx=None):
    _check_inplace(mask, np.float32, mask.shape)
    c = _Args(ctx).ctx_for(mask)
    h, w = mask.shape
    check(c._lib.dm_enlarge_mask(c.handle, _ptr(mask), h, w, int(ix), int(iy)))
    return mask


# test_opticalflow.lua:143-216
def radial(geometry, flow, mh=None, mw=None, ctx=None):
    args = _Args(ctx)
    fptr, f = args.inp(flow)
    c = args.ctx_for(f)
    _, h, w = f.shape
    mh = h / 2 if mh is None else mh
    mw = w / 2 if mw is None else mw
    rptr, ret = args.out((h, w), np.float32, like=f)
    cptr, conf = args.out((h, w), np.float32, like=f)
    check(c._lib.dm_radial_depth(c.handle, fptr, h, w, float(mh), float(mw), float(geometry.wImg / 2),
                                 rptr, cptr))
    return ret, conf


# ardrone/ardrone_api.cpp:99-140
def computeDepthMapFromFlow(xflow, mask, imu_translation_x, ctx=None):
    args = _Args(ctx)
    xptr, x = args.inp(xflow)
    mptr, m = args.inp(mask)
    c = args.ctx_for(x, m)
    h, w = x.shape
    dptr, depth = args.out((h, w), np.float32, like=x)
    cptr, conf = args.out((h, w), np.float32, like=x)
    check(c._lib.dm_depth_from_xflow(c.handle, xptr, mptr, h, w, float(imu_translation_x), dptr, cptr))
    return depth, conf
