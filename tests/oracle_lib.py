"""ctypes bindings for the CPU oracle (oracle/libdm_oracle.so) and, when built, the
reference's own native sources (oracle/_ref/libdm_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_ORACLE_SO = os.path.join(ORACLE_DIR, "libdm_oracle.so")
_REF_SO = os.path.join(ORACLE_DIR, "_ref", "libdm_ref.so")

c_fp = C.POINTER(C.c_float)
c_lp = C.POINTER(C.c_int64)
c_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_ORACLE_SO):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    return _ORACLE_SO


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(c_fp)


def _l(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(c_lp)


def _ints(v):
    arr = (C.c_int * len(v))(*[int(x) for x in v])
    return arr


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_ORACLE_SO)
        _lib.orc_extract_output.restype = C.c_int64
        _lib.orc_extract_output_marginalized.restype = C.c_int64
        _lib.orc_yx2x_multi.restype = C.c_int64
        _lib.orc_yx2x_multi.argtypes = [C.c_int, C.c_int, c_ip, C.c_int, C.c_double, C.c_double]
        _lib.orc_get_rmax.restype = C.c_double
        _lib.orc_get_rmax.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
    return _lib


_ref = None


def ref():
    """The reference's own code (oracle/_ref); None when it was never built."""
    global _ref
    if _ref is None and os.path.exists(_REF_SO):
        _ref = C.CDLL(_REF_SO)
    return _ref


# ---------------------------------------------------------------- matching
def spatial_matching(in1, in2, maxh, maxw, nthreads=0):
    in1, p1 = _f(in1)
    in2, p2 = _f(in2)
    Cn, H1, W1 = in1.shape
    _, H2, W2 = in2.shape
    assert H2 >= H1 + maxh - 1 and W2 >= W1 + maxw - 1
    out = np.empty((H1, W1, maxh, maxw), np.float32)
    lib().orc_spatial_matching(p1, p2, Cn, H1, W1, H2, W2, maxh, maxw, out.ctypes.data_as(c_fp),
                               int(nthreads))
    return out


def radial_matching(in1, in2, hwin, nthreads=0):
    in1, p1 = _f(in1)
    in2, p2 = _f(in2)
    Cn, H1, W = in1.shape
    H2 = in2.shape[1]
    out = np.empty((H1, W, hwin), np.float32)
    lib().orc_radial_matching(p1, p2, Cn, H1, W, H2, hwin, out.ctypes.data_as(c_fp), int(nthreads))
    return out


def neg_softmax(vol, exp_mode=0, nthreads=0, K=None):
    """Softmax over the window: the last two dims of a (..,H1,W1,maxh,maxw) volume, else the last."""
    vol, pv = _f(vol)
    if K is None:
        K = vol.shape[-1] * vol.shape[-2] if vol.ndim >= 4 else vol.shape[-1]
    rows = vol.size // K
    out = np.empty_like(vol)
    lib().orc_neg_softmax(pv, C.c_int64(rows), K, exp_mode, out.ctypes.data_as(c_fp), int(nthreads))
    return out


def argmax_tie(prob, K, middle):
    prob, pp = _f(prob)
    rows = prob.size // K
    idx = np.empty(rows, np.int64)
    mx = np.empty(rows, np.float32)
    lib().orc_argmax_tie(pp, C.c_int64(rows), K, middle, idx.ctypes.data_as(c_lp),
                         mx.ctypes.data_as(c_fp))
    return idx, mx


def argmin_tie(vol, K, middle):
    vol, pv = _f(vol)
    rows = vol.size // K
    idx = np.empty(rows, np.int64)
    mn = np.empty(rows, np.float32)
    lib().orc_argmin_tie(pv, C.c_int64(rows), K, middle, idx.ctypes.data_as(c_lp),
                         mn.ctypes.data_as(c_fp))
    return idx, mn


def top2_relgap(prob, K):
    prob, pp = _f(prob)
    rows = prob.size // K
    g = np.empty(rows, np.float32)
    lib().orc_top2_relgap(pp, C.c_int64(rows), K, g.ctypes.data_as(c_fp))
    return g


def soft_mean(prob, maxh, maxw):
    prob, pp = _f(prob)
    rows = prob.size // (maxh * maxw)
    ym = np.empty(rows, np.float32)
    xm = np.empty(rows, np.float32)
    lib().orc_soft_mean(pp, C.c_int64(rows), maxh, maxw, ym.ctypes.data_as(c_fp),
                        xm.ctypes.data_as(c_fp))
    return ym, xm


def marginal_x(prob, maxh, maxw):
    prob, pp = _f(prob)
    rows = prob.size // (maxh * maxw)
    pm = np.empty((rows, maxh), np.float32)
    lib().orc_marginal_x(pp, C.c_int64(rows), maxh, maxw, pm.ctypes.data_as(c_fp))
    return pm


def flow_canvas(idx, h1, w1, maxh, maxw, hImg, wImg):
    idx, pi = _l(idx)
    full = np.empty((2, hImg, wImg), np.float32)
    lib().orc_flow_canvas(pi, h1, w1, maxh, maxw, hImg, wImg, full.ctypes.data_as(c_fp))
    return full


# ----------------------------------------------------------------- extract
def extract_output(inp, threshold, ret=None, scores=None, which="oracle"):
    """Returns (ret, scores, written).  ret/scores start as the given arrays (the
    reference leaves untouched pixels alone) or zeros."""
    inp, pin = _f(inp)
    h, w, n = inp.shape
    ret = np.zeros((h, w), np.int64) if ret is None else np.ascontiguousarray(ret, np.int64).copy()
    scores = (np.zeros((h, w), np.float32) if scores is None
              else np.ascontiguousarray(scores, np.float32).copy())
    if which == "oracle":
        written = lib().orc_extract_output(pin, h, w, n, C.c_double(threshold),
                                           ret.ctypes.data_as(c_lp), scores.ctypes.data_as(c_fp))
    else:
        r = ref()
        assert r is not None, "oracle/_ref not built"
        r.ref_extract_output(pin, C.c_long(h), C.c_long(w), C.c_long(n), C.c_double(threshold),
                             ret.ctypes.data_as(c_lp), scores.ctypes.data_as(c_fp))
        written = None
    return ret, scores, written


def extract_output_marginalized(inp, threshold, threshold_acc, ret=None, which="oracle"):
    inp, pin = _f(inp)
    h, w, n = inp.shape
    ret = np.zeros((h, w), np.int64) if ret is None else np.ascontiguousarray(ret, np.int64).copy()
    gd = np.full((h, w), 7, np.int64)
    if which == "oracle":
        lib().orc_extract_output_marginalized(pin, h, w, n, C.c_double(threshold),
                                              C.c_double(threshold_acc), ret.ctypes.data_as(c_lp),
                                              gd.ctypes.data_as(c_lp))
    else:
        r = ref()
        assert r is not None, "oracle/_ref not built"
        r.ref_extract_output_marginalized(pin, C.c_long(h), C.c_long(w), C.c_long(n),
                                          C.c_double(threshold), C.c_double(threshold_acc),
                                          ret.ctypes.data_as(c_lp), gd.ctypes.data_as(c_lp))
    return ret, gd


# -------------------------------------------------------------- multiscale
def yx2x_multi(maxh, maxw, ratios, y, x):
    return int(lib().orc_yx2x_multi(maxh, maxw, _ints(ratios), len(ratios), float(y), float(x)))


def x2yx_multi_number(maxh, maxw, ratios, x):
    oy = C.c_int64(0)
    ox = C.c_int64(0)
    rc = lib().orc_x2yx_multi_number(maxh, maxw, _ints(ratios), len(ratios), C.c_int64(int(x)),
                                     C.byref(oy), C.byref(ox))
    return rc, oy.value, ox.value


def x2yx_multi2_bugcompat(xim, maxh, maxw, ratios, which="oracle", fill=-777):
    xim, px = _l(xim)
    h, w = xim.shape
    retx = np.full((h, w), fill, np.int64)
    rety = np.full((h, w), fill, np.int64)
    if which == "oracle":
        lib().orc_x2yx_multi2_bugcompat(px, h, w, maxh, maxw, _ints(ratios), len(ratios),
                                        retx.ctypes.data_as(c_lp), rety.ctypes.data_as(c_lp))
    else:
        r = ref()
        assert r is not None, "oracle/_ref not built"
        rat = (C.c_double * len(ratios))(*[float(v) for v in ratios])
        r.ref_x2yx_multi2(px, C.c_long(h), C.c_long(w), maxh, maxw, rat, len(ratios),
                          retx.ctypes.data_as(c_lp), rety.ctypes.data_as(c_lp))
    return rety, retx


def multiscale_length(maxh, maxw, ratios):
    return int(lib().orc_multiscale_length(maxh, maxw, _ints(ratios), len(ratios)))


def cascade_add(inp, ratios):
    """inp: (nratios, rows, Kh, Kw)"""
    inp, pin = _f(inp)
    n, rows, Kh, Kw = inp.shape
    out = np.empty_like(inp)
    lib().orc_cascade_add(pin, C.c_int64(rows), Kh, Kw, _ints(ratios), n, out.ctypes.data_as(c_fp))
    return out


def ring_join(casc, ratios):
    casc, pc = _f(casc)
    n, rows, Kh, Kw = casc.shape
    L = multiscale_length(Kh, Kw, ratios)
    out = np.empty((rows, L), np.float32)
    lib().orc_ring_join(pc, C.c_int64(rows), Kh, Kw, _ints(ratios), n, out.ctypes.data_as(c_fp))
    return out


def downsample_avg(img, r):
    img, pi = _f(img)
    Cn, H, W = img.shape
    out = np.empty((Cn, H // r, W // r), np.float32)
    lib().orc_downsample_avg(pi, Cn, H, W, r, out.ctypes.data_as(c_fp))
    return out


def upsample_nearest_rows(v, r):
    v, pv = _f(v)
    h, w, K = v.shape
    out = np.empty((h * r, w * r, K), np.float32)
    lib().orc_upsample_nearest_rows(pv, h, w, K, r, out.ctypes.data_as(c_fp))
    return out


# ------------------------------------------------------------------ radial
def c2p_mask(wdst, hdst, xc, yc, lpad, rpad, rmax, alpha=1.0):
    m = np.empty((2, hdst, wdst + lpad + rpad), np.float32)
    lib().orc_c2p_mask(wdst, hdst, C.c_double(xc), C.c_double(yc), lpad, rpad, C.c_double(rmax),
                       C.c_double(alpha), m.ctypes.data_as(c_fp))
    return m


def p2c_mask(wsrc, hsrc, wdst, hdst, xc, yc, rmax, alpha=1.0):
    m = np.empty((2, hdst, wdst), np.float32)
    lib().orc_p2c_mask(wsrc, hsrc, wdst, hdst, C.c_double(xc), C.c_double(yc), C.c_double(rmax),
                       C.c_double(alpha), m.ctypes.data_as(c_fp))
    return m


def get_rmax(h, w, ex, ey):
    return float(lib().orc_get_rmax(h, w, ex, ey))


def warp_bilinear(src, field):
    src, ps = _f(src)
    field, pf = _f(field)
    Cn, hs, ws = src.shape
    _, hd, wd = field.shape
    dst = np.empty((Cn, hd, wd), np.float32)
    lib().orc_warp_bilinear(ps, Cn, hs, ws, pf, hd, wd, dst.ctypes.data_as(c_fp))
    return dst


def flow2depth(flow, xc, yc, infty):
    flow, pf = _f(flow)
    h, w = flow.shape
    depth = np.empty((h, w), np.float32)
    confs = np.empty((h, w), np.float32)
    lib().orc_flow2depth(pf, h, w, C.c_float(xc), C.c_float(yc), C.c_float(infty),
                         depth.ctypes.data_as(c_fp), confs.ctypes.data_as(c_fp))
    return depth, confs


# ------------------------------------------------------------- postprocess
def post_process_image(inp, mask, k, method, which="oracle"):
    """postProcessImage(input [2,h,w], mask [h,w], winsize, 'med'|'max')."""
    inp, pi = _f(inp)
    mask, pm = _f(mask)
    _, h, w = inp.shape
    out = np.zeros((2, h, w), np.float32)
    if which == "oracle":
        lib().orc_post_process_image(pi, pm, h, w, k, 1 if method == "max" else 0,
                                     out.ctypes.data_as(c_fp))
        return out
    r = ref()
    assert r is not None, "oracle/_ref not built"
    if method == "max":  # the Lua around the inline C (opticalflow_model.lua:435-439)
        rr = np.floor(inp + np.float32(0.5)).astype(np.float32)
        m = rr.min()
        rr = np.ascontiguousarray(rr - m)
        r.ref_pp_filter(1, rr.ctypes.data_as(c_fp), pm, k, C.c_long(h), C.c_long(w), out.ctypes.data_as(c_fp))
        return out + m
    r.ref_pp_filter(0, pi, pm, k, C.c_long(h), C.c_long(w), out.ctypes.data_as(c_fp))
    return out


def enlarge_mask(mask, ix, iy, which="oracle"):
    m = np.ascontiguousarray(mask, np.float32).copy()
    h, w = m.shape
    if which == "oracle":
        lib().orc_enlarge_mask(m.ctypes.data_as(c_fp), h, w, ix, iy)
    else:
        ref().ref_enlarge_mask(m.ctypes.data_as(c_fp), C.c_long(h), C.c_long(w), ix, iy)
    return m


def radial_depth(flow, mh, mw, infty, which="oracle"):
    flow, pf = _f(flow)
    _, h, w = flow.shape
    ret = np.zeros((h, w), np.float32)
    conf = np.zeros((h, w), np.float32)
    if which == "oracle":
        lib().orc_radial_depth(pf, h, w, C.c_float(mh), C.c_float(mw), C.c_float(infty),
                               ret.ctypes.data_as(c_fp), conf.ctypes.data_as(c_fp))
    else:
        ref().ref_radial_depth(pf, C.c_long(h), C.c_long(w), C.c_double(mh), C.c_double(mw),
                               ret.ctypes.data_as(c_fp), conf.ctypes.data_as(c_fp), C.c_double(infty))
    return ret, conf


def depth_from_xflow(xflow, mask, m, which="oracle"):
    """computeDepthMapFromFlow (ardrone/ardrone_api.cpp:99-140); which="ref" runs the reference's own
    statements compiled into oracle/_ref (depth is unspecified there where the confidence is 0)."""
    xflow, px = _f(xflow)
    mask, pm = _f(mask)
    h, w = xflow.shape
    depth = np.empty((h, w), np.float32)
    conf = np.empty((h, w), np.float32)
    if which == "oracle":
        fn = lib().orc_depth_from_xflow
    else:
        assert ref() is not None, "oracle/_ref not built"
        fn = ref().ref_depth_from_xflow
    fn(px, pm, h, w, C.c_float(m), depth.ctypes.data_as(c_fp), conf.ctypes.data_as(c_fp))
    return depth, conf


def ref_c2p_mask(wdst, hdst, xc, yc, rmax, alpha=1.0):
    """getC2PMask's inline C through oracle/_ref (un-padded), with the Lua-side constants."""
    import math
    m = np.empty((2, hdst, wdst), np.float32)
    ref().ref_c2p_mask(m.ctypes.data_as(c_fp), C.c_long(hdst), C.c_long(wdst), C.c_double(xc),
                       C.c_double(yc), C.c_double(rmax / (hdst ** alpha)),
                       C.c_double(2 * math.pi / wdst), C.c_double(alpha))
    return m


def ref_p2c_mask(wsrc, hsrc, wdst, hdst, xc, yc, rmax, alpha=1.0):
    import math
    m = np.empty((2, hdst, wdst), np.float32)
    pi2 = 2 * math.pi
    ref().ref_p2c_mask(m.ctypes.data_as(c_fp), C.c_long(hdst), C.c_long(wdst), C.c_double(xc),
                       C.c_double(yc), C.c_double(wsrc / pi2), C.c_double(hsrc / (rmax ** (1.0 / alpha))),
                       C.c_double(pi2), C.c_double(1.0 / alpha))
    return m


def ref_flow2depth(flow, xc, yc, infty):
    flow, pf = _f(flow)
    h, w = flow.shape
    depth = np.zeros((h, w), np.float32)   # `ret` starts at zero, `confs` at one (display.lua:12-13)
    confs = np.ones((h, w), np.float32)
    ref().ref_flow2depth(pf, C.c_long(h), C.c_long(w), depth.ctypes.data_as(c_fp),
                         confs.ctypes.data_as(c_fp), C.c_double(xc), C.c_double(yc), C.c_double(infty))
    return depth, confs


# ------------------------------------------------------------- filter (feature extractor)
def conv_layer(inp, weight, bias, conn=None, pads=(0, 0, 0, 0), tanh=False, nthreads=0):
    """One getFilter layer.  inp [n_in,h,w]; weight [n_out,n_in,kh,kw] (conn None) or
    [n_conn,kh,kw] with conn [n_conn,2] 1-based (from,to); pads = (l, r, t, b)."""
    inp, pi = _f(inp)
    weight, pw = _f(weight)
    bias, pb = _f(bias)
    n_in, h, w = inp.shape
    n_out = bias.shape[0]
    kh, kw = weight.shape[-2:]
    if conn is None:
        assert weight.shape[:2] == (n_out, n_in)
        pc, nc = None, 0
    else:
        conn = np.ascontiguousarray(conn, np.int32)
        assert conn.shape == (weight.shape[0], 2)
        pc, nc = conn.ctypes.data_as(C.c_void_p), conn.shape[0]
    pl, pr, pt, pbm = pads
    ho, wo = h + pt + pbm - kh + 1, w + pl + pr - kw + 1
    out = np.empty((n_out, ho, wo), np.float32)
    lib().orc_conv_layer(pi, n_in, h, w, pw, pb, n_out, kh, kw, pc, nc, pl, pr, pt, pbm, int(bool(tanh)),
                         out.ctypes.data_as(c_fp), int(nthreads))
    return out


def filter_forward(inp, layers, pads=(0, 0, 0, 0)):
    """layers: list of dicts(weight, bias, conn=None, tanh=False); padding before the first."""
    x = inp
    for i, L in enumerate(layers):
        x = conv_layer(x, L["weight"], L["bias"], L.get("conn"), pads if i == 0 else (0, 0, 0, 0),
                       L.get("tanh", False))
    return x


def warp_homography(src, hmat, hd=None, wd=None):
    """removeEgoMotion as a homography gather: returns (warped [C,hd,wd], mask [hd,wd])."""
    src, ps = _f(src)
    Cn, hs, ws = src.shape
    hd, wd = hd or hs, wd or ws
    hm = np.ascontiguousarray(hmat, np.float64).reshape(9)
    dst = np.empty((Cn, hd, wd), np.float32)
    mask = np.empty((hd, wd), np.float32)
    lib().orc_warp_homography(ps, Cn, hs, ws, hm.ctypes.data_as(C.POINTER(C.c_double)), hd, wd,
                              dst.ctypes.data_as(c_fp), mask.ctypes.data_as(c_fp))
    return dst, mask
