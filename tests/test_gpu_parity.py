"""Parity of the CUDA path (through the C ABI) with the CPU oracle on identical seeded
inputs.  Bars (BASELINE.json north_star): integer indices bit-exact, except pixels whose
top-2 softmax scores differ by < 1e-5 relative (counted and reported); fp32 scores within
1e-4 relative.  With DM_FLAG_EXACT_SSD the SSD itself is bit-exact."""
import math

import numpy as np
import pytest

from synth import make_pair

pytestmark = pytest.mark.gpu

RTOL = 1e-4          # fp32 scores (north_star)
NEAR_TIE = 1e-5      # relative top-2 gap below which an index may differ (north_star)


def oracle_outputs(oracle, in1, in2, maxh, maxw, thr=0.11):
    K = maxh * maxw
    vol = oracle.spatial_matching(in1, in2, maxh, maxw)
    prob = oracle.neg_softmax(vol)
    cy, cx = math.ceil(maxh / 2), math.ceil(maxw / 2)
    middle = (cy - 1) * maxw + cx
    idx, pmax = oracle.argmax_tie(prob, K, middle)
    shp = in1.shape[1:]
    ret, sc, _ = oracle.extract_output(prob.reshape(shp + (K,)), thr)
    ym, xm = oracle.soft_mean(prob, maxh, maxw)
    gap = oracle.top2_relgap(prob, K)
    return dict(vol=vol, prob=prob, index=idx.reshape(shp), pmax=pmax.reshape(shp),
                min_ssd=vol.reshape(shp + (K,)).min(-1), index_thr=ret, score_thr=sc,
                soft_y=ym.reshape(shp), soft_x=xm.reshape(shp), gap=gap.reshape(shp), middle=middle)


def check_fused(oracle, got, want, exact, thr=0.11):
    tie = want["gap"] < NEAR_TIE
    bad = (got["index"] != want["index"]) & ~tie
    assert bad.sum() == 0, "%d index mismatches outside near-ties (near-ties: %d)" % (bad.sum(), tie.sum())
    if exact:
        np.testing.assert_array_equal(got["min_ssd"], want["min_ssd"])
    else:
        np.testing.assert_allclose(got["min_ssd"], want["min_ssd"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(got["pmax"], want["pmax"], rtol=RTOL)
    np.testing.assert_allclose(got["soft_yx"][0], want["soft_y"], rtol=RTOL)
    np.testing.assert_allclose(got["soft_yx"][1], want["soft_x"], rtol=RTOL)
    # thresholded extraction: a probability within 1e-5 relative of the threshold may flip
    prob = want["prob"].reshape(want["index"].shape + (-1,))
    edge = (np.abs(prob - thr) < thr * 1e-5 * 4).any(-1) | tie
    ok = ~edge
    np.testing.assert_array_equal(got["index_thr"][ok], want["index_thr"][ok])
    np.testing.assert_allclose(got["score_thr"][ok], want["score_thr"][ok], rtol=RTOL, atol=1e-7)
    return int(tie.sum())


CASES = [
    # C, H2, W2, maxh, maxw
    (10, 40, 60, 17, 17),
    (10, 50, 170, 33, 33),    # more than one tile in x, tail block of 9
    (3, 21, 37, 5, 7),        # ragged, W2 not a multiple of 4 (repack path), CT=4
    (1, 12, 20, 1, 1),        # degenerate 1x1 window
    (4, 30, 45, 8, 8),
    (10, 30, 50, 16, 16),
    (16, 26, 44, 9, 10),      # CT=16, tail block of 2
    (7, 40, 90, 15, 65),      # widest window
    (20, 24, 30, 6, 6),       # > 16 channels: untiled kernel
]


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("case", CASES)
def test_fused_match_extract_vs_oracle(dm, oracle, case, exact):
    C, H2, W2, maxh, maxw = case
    in1, in2, _ = make_pair(C, H2, W2, maxh, maxw, seed=11, noise=0.3)
    want = oracle_outputs(oracle, in1, in2, maxh, maxw)
    got = dm.match_extract(in1, in2, maxh, maxw, exact=exact,
                           want=("index", "min_ssd", "pmax", "index_thr", "score_thr", "soft_yx",
                                 "n_untouched"))
    check_fused(oracle, got, want, exact)
    assert int(got["n_untouched"]) >= int((want["index_thr"] == 0).sum()) - 4


@pytest.mark.parametrize("noise", [0.0, 0.05])
def test_planted_flow_is_recovered(dm, noise):
    """T2 on the GPU path: the planted integer flow is the known answer."""
    maxh = maxw = 17
    in1, in2, flow = make_pair(10, 80, 120, maxh, maxw, seed=3, noise=noise)
    got = dm.match_extract(in1, in2, maxh, maxw, canvas=(80, 120), want=("index",))
    full = got["flow_full"]
    hoff, woff = (80 - in1.shape[1]) // 2, (120 - in1.shape[2]) // 2
    np.testing.assert_array_equal(full[0, hoff:hoff + in1.shape[1], woff:woff + in1.shape[2]], flow[0])
    np.testing.assert_array_equal(full[1, hoff:hoff + in1.shape[1], woff:woff + in1.shape[2]], flow[1])
    assert full[:, :hoff].sum() == 0 and full[:, hoff + in1.shape[1]:].sum() == 0


def test_zero_flow_tie_rule(dm, oracle):
    """Constant frames: every window entry ties, the middle index must win
    (opticalflow_model.lua:157-159), and softmax is uniform."""
    maxh, maxw = 5, 9
    in2 = np.ones((3, 20, 40), np.float32)
    in1 = np.ones((3, 16, 32), np.float32)
    got = dm.match_extract(in1, in2, maxh, maxw, want=("index", "pmax", "index_thr", "score_thr"))
    want = oracle_outputs(oracle, in1, in2, maxh, maxw)
    np.testing.assert_array_equal(got["index"], want["index"])
    assert (got["index"] == want["middle"]).all()
    np.testing.assert_allclose(got["pmax"], 1.0 / 45, rtol=1e-5)
    assert (got["index_thr"] == 0).all() and (got["score_thr"] == 0).all()
    got = dm.match_extract(in1, in2, maxh, maxw, tie_middle=False, want=("index",))
    assert (got["index"] == 1).all()


def test_small_window_all_qualify_scan_order(dm, oracle):
    """3x3 window of equal SSDs: p = 1/9 > 0.11 for all nine, extractOutput keeps the FIRST 8."""
    in2 = np.zeros((2, 10, 12), np.float32)
    in1 = np.zeros((2, 8, 10), np.float32)
    want = oracle_outputs(oracle, in1, in2, 3, 3)
    got = dm.match_extract(in1, in2, 3, 3, want=("index_thr", "score_thr"))
    np.testing.assert_array_equal(got["index_thr"], want["index_thr"])
    np.testing.assert_allclose(got["score_thr"], want["score_thr"], rtol=1e-6)


def test_batched_strided_and_device_inputs(dm, oracle):
    """N > 1, in1 given as the strided crop prepareInput returns, torch CUDA tensors in and out."""
    import torch
    maxh, maxw, C = 9, 17, 10
    frames = []
    for n in range(3):
        in1, in2, _ = make_pair(C, 40, 72, maxh, maxw, seed=20 + n, noise=0.2)
        full1 = np.random.default_rng(n).standard_normal((C, 40, 72)).astype(np.float32)
        oy, ox = math.ceil(maxh / 2) - 1, math.ceil(maxw / 2) - 1
        full1[:, oy:oy + in1.shape[1], ox:ox + in1.shape[2]] = in1
        frames.append((full1, in2, in1))
    f1 = np.stack([f[0] for f in frames])
    f2 = np.stack([f[1] for f in frames])
    g = dm.Geometry(maxh=maxh, maxw=maxw)
    oy, ox = math.ceil(maxh / 2) - 1, math.ceil(maxw / 2) - 1
    H1, W1 = 40 - maxh + 1, 72 - maxw + 1
    view = f1[:, :, oy:oy + H1, ox:ox + W1]
    host = dm.match_extract(view, f2, maxh, maxw, exact=True, want=("index", "min_ssd"))
    t1 = torch.from_numpy(f1).cuda()[:, :, oy:oy + H1, ox:ox + W1]
    t2 = torch.from_numpy(f2).cuda()
    dev = dm.match_extract(t1, t2, maxh, maxw, exact=True, want=("index", "min_ssd"))
    torch.cuda.synchronize()
    for n in range(3):
        want = oracle_outputs(oracle, frames[n][2], frames[n][1], maxh, maxw)
        tie = want["gap"] < NEAR_TIE
        assert ((host["index"][n] != want["index"]) & ~tie).sum() == 0
        np.testing.assert_array_equal(host["min_ssd"][n], want["min_ssd"])
        np.testing.assert_array_equal(dev["index"][n].cpu().numpy(), host["index"][n])
        np.testing.assert_array_equal(dev["min_ssd"][n].cpu().numpy(), host["min_ssd"][n])


@pytest.mark.parametrize("case", [(10, 30, 150, 9, 17), (3, 19, 23, 4, 5), (20, 16, 20, 3, 4)])
def test_volume_ssd_and_softmax_vs_oracle(dm, oracle, case):
    C, H2, W2, maxh, maxw = case
    in1, in2, _ = make_pair(C, H2, W2, maxh, maxw, seed=2, noise=0.5)
    vol = oracle.spatial_matching(in1, in2, maxh, maxw)
    got = dm.nn.SpatialMatching(maxh, maxw, False, exact=True).forward([in1, in2])
    np.testing.assert_array_equal(got, vol)
    got = dm.nn.SpatialMatching(maxh, maxw).forward([in1, in2])
    np.testing.assert_allclose(got, vol, rtol=RTOL, atol=1e-6)
    prob = oracle.neg_softmax(vol)
    gotp = dm.match_volume(in1, in2, maxh, maxw, softmax=True)
    np.testing.assert_allclose(gotp, prob, rtol=RTOL, atol=1e-9)
    # stand-alone softmax on a volume in memory
    from depthmatch import api
    a = api._Args()
    K = maxh * maxw
    out = np.empty_like(vol)
    c = dm.default_context()
    api.check(c._lib.dm_neg_softmax(c.handle, vol.ctypes.data, vol.size // K, K, out.ctypes.data))
    np.testing.assert_allclose(out, prob, rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("case", [(10, 30, 152, 9, 17), (4, 12, 100, 5, 33), (10, 20, 71, 8, 8), (16, 40, 64, 33, 33),
                                  (3, 26, 44, 17, 9), (10, 21, 40, 4, 25), (4, 9, 12, 9, 9), (10, 5, 20, 4, 17),
                                  (1, 3, 10, 1, 3), (10, 34, 36, 33, 33)])
def test_volume_strip_kernel_vs_oracle_and_tiled_kernel(dm, oracle, case):
    """W1 % 4 == 0: the strip kernel (whole pixel streams, bulk copies; match_volume_px.cuh).  Against the
    oracle, and bit for bit against the tiled kernel (option volume_kernel = 1)."""
    C, H2, W2, maxh, maxw = case
    in1, in2, _ = make_pair(C, H2, W2, maxh, maxw, seed=5, noise=0.5)
    assert in1.shape[2] % 4 == 0
    vol = oracle.spatial_matching(in1, in2, maxh, maxw)
    prob = oracle.neg_softmax(vol)
    ctx = dm.default_context()
    res = {}
    for kern in (2, 1):   # 2 = the strip kernel also for windows it would leave to the tiled one
        ctx.set_option("volume_kernel", kern)
        res[kern & 1] = (dm.match_volume(in1, in2, maxh, maxw, exact=True), dm.match_volume(in1, in2, maxh, maxw),
                     dm.match_volume(in1, in2, maxh, maxw, softmax=True))
    ctx.set_option("volume_kernel", 0)
    np.testing.assert_array_equal(res[0][0], vol)
    np.testing.assert_allclose(res[0][1], vol, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(res[0][2], prob, rtol=RTOL, atol=1e-9)
    # same block functions in both kernels: the SSD volumes agree bit for bit; the strip kernel's soft-max
    # takes min and sum in its staging buffer (another summation order than the statistics sweep)
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][1], res[1][1])
    np.testing.assert_allclose(res[0][2], res[1][2], rtol=2e-5, atol=1e-12)


def test_volume_strip_kernel_writes_only_its_output(dm):
    """Bulk copies of whole pixel streams: guard words before and after the caller's buffer stay untouched
    (partial last strip, several pairs, both volume modes) and the result equals the ordinary call's."""
    import ctypes
    import torch
    from depthmatch import _lib as dml
    from depthmatch import api
    rng = np.random.default_rng(21)
    N, C, maxh, maxw = 2, 10, 9, 17
    H1, W1 = 20, 136
    H2, W2 = H1 + maxh - 1, W1 + maxw - 1
    t2 = torch.from_numpy(rng.standard_normal((N, C, H2, W2), dtype=np.float32)).cuda()
    t1 = (t2[:, :, 4:4 + H1, 8:8 + W1] + 0.2).contiguous()
    ctx = dm.default_context()
    ctx.set_option("volume_kernel", 2)
    K, G = maxh * maxw, 4096
    n_out = N * H1 * W1 * K
    pr = dml.dm_pair()
    pr.in1, pr.in2 = t1.data_ptr(), t2.data_ptr()
    pr.n_pairs, pr.channels, pr.h1, pr.w1, pr.h2, pr.w2 = N, C, H1, W1, H2, W2
    for mode, softmax in ((dml.DM_VOLUME_SSD, False), (dml.DM_VOLUME_NEG_SOFTMAX, True)):
        buf = torch.full((G + n_out + G,), -12345.0, device="cuda")
        api.check(ctx._lib.dm_match_volume(ctx.handle, ctypes.byref(pr), maxh, maxw, mode,
                                           ctypes.c_void_p(buf.data_ptr() + 4 * G)))
        torch.cuda.synchronize()
        assert bool((buf[:G] == -12345.0).all()) and bool((buf[G + n_out:] == -12345.0).all())
        want = dm.match_volume(t1, t2, maxh, maxw, softmax=softmax)
        assert torch.equal(buf[G:G + n_out].view(want.shape), want)
    ctx.set_option("volume_kernel", 0)


def test_volume_strip_kernel_many_units_per_cta(dm, oracle):
    """More units than SMs (barrier phases carried across units), a partial last strip and row bands."""
    import torch
    rng = np.random.default_rng(11)
    N, C, maxh, maxw = 5, 10, 9, 9
    H1, W1 = 45, 168
    in2 = rng.standard_normal((N, C, H1 + maxh - 1, W1 + maxw - 1), dtype=np.float32)
    in1 = (in2[:, :, 4:4 + H1, 3:3 + W1] + 0.3 * rng.standard_normal((N, C, H1, W1), dtype=np.float32)).copy()
    t1, t2 = torch.from_numpy(in1).cuda(), torch.from_numpy(in2).cuda()
    dm.default_context().set_option("volume_kernel", 2)
    got = dm.match_volume(t1, t2, maxh, maxw, exact=True).cpu().numpy()
    gotp = dm.match_volume(t1, t2, maxh, maxw, softmax=True).cpu().numpy()
    dm.default_context().set_option("volume_kernel", 0)
    for n in range(N):
        vol = oracle.spatial_matching(in1[n], in2[n], maxh, maxw)
        np.testing.assert_array_equal(got[n], vol)
        np.testing.assert_allclose(gotp[n], oracle.neg_softmax(vol), rtol=RTOL, atol=1e-9)


@pytest.mark.parametrize("case", [(10, 30, 150, 9, 17, 0.21, 0.5), (3, 19, 23, 4, 5, 0.21, 0.05), (10, 24, 40, 7, 7, 0.15, 0.1),
                                  (20, 16, 20, 3, 4, 0.21, 0.02)])
def test_raw_ssd_thresholded_extraction_vs_oracle(dm, oracle, case):
    """extractOutput(volume of raw SSDs, 0.21) as the ground-truth generators call it
    (radial/radial_opticalflow_groundtruth.lua:105, version2/groundtruth.lua:103), fused: no volume.  Oracle =
    the reference's extract_output.cpp (restated, and the compiled original where built) on the oracle's volume.
    Small noise keeps some SSDs under the threshold, so the scan has to pass over them; bit-exact."""
    C, H2, W2, maxh, maxw, thr, noise = case
    in1, in2, _ = make_pair(C, H2, W2, maxh, maxw, seed=9, noise=noise)
    vol = oracle.spatial_matching(in1, in2, maxh, maxw)
    K = maxh * maxw
    want_ret, want_sc, written = oracle.extract_output(vol.reshape(vol.shape[0], vol.shape[1], K), thr)
    ret, sc, untouched = dm.match_extract_raw_ssd(in1, in2, maxh, maxw, thr)
    np.testing.assert_array_equal(ret, want_ret)
    np.testing.assert_array_equal(sc, want_sc)
    assert untouched == int((want_ret == 0).sum())
    below = (vol.reshape(-1, K)[:, :8] <= thr).any()
    if noise <= 0.05:
        assert below      # the case really has entries the scan must skip
    # batched device call
    import torch
    t1 = torch.from_numpy(np.stack([in1, in1])).cuda()
    t2 = torch.from_numpy(np.stack([in2, in2])).cuda()
    r2, s2, u2 = dm.match_extract_raw_ssd(t1, t2, maxh, maxw, thr)
    np.testing.assert_array_equal(r2[1].cpu().numpy(), want_ret)
    np.testing.assert_array_equal(s2[0].cpu().numpy(), want_sc)
    assert list(u2) == [untouched, untouched]


def test_module_level_process_output_vs_oracle(dm, oracle):
    """getModel(prefiltered) -> forward -> processOutput for 'max', 'max'+threshold and 'mean'."""
    maxh, maxw = 9, 9
    in1, in2, _ = make_pair(10, 40, 56, maxh, maxw, seed=8, noise=0.3)
    want = oracle_outputs(oracle, in1, in2, maxh, maxw)
    g = dm.Geometry(maxh=maxh, maxw=maxw, hImg=40, wImg=56, output_extraction_method="max")
    model = dm.getModel(g, True, True)
    prob = model.forward([in1, in2])
    np.testing.assert_allclose(prob.reshape(want["prob"].shape), want["prob"], rtol=RTOL, atol=1e-9)
    # feed the ORACLE's volume so that the stand-alone kernels are checked bit for bit
    oprob = want["prob"].reshape(in1.shape[1:] + (maxh * maxw,))
    out = dm.processOutput(g, oprob, True, None)
    np.testing.assert_array_equal(out["index"], want["index"])
    full = oracle.flow_canvas(want["index"], in1.shape[1], in1.shape[2], maxh, maxw, 40, 56)
    np.testing.assert_array_equal(out["full"], full)
    out = dm.processOutput(g, oprob, True, 2)
    np.testing.assert_array_equal(out["index"], want["index_thr"])
    np.testing.assert_array_equal(out["confidences"], want["score_thr"] > 2)
    g.output_extraction_method = "mean"
    out = dm.processOutput(g, oprob, True, None)
    cy = math.ceil(maxh / 2)
    np.testing.assert_array_equal(out["y"] + cy, want["soft_y"])
    np.testing.assert_array_equal(out["x"] + cy, want["soft_x"])
    pm = oracle.marginal_x(want["prob"], maxh, maxw).reshape(in1.shape[1:] + (maxh,))
    _, sc, _ = oracle.extract_output(pm, 0.11)
    np.testing.assert_array_equal(out["confidences"], sc > 0)
    # and the fused module gives the same answers without the volume
    g.output_extraction_method = "max"
    fused = dm.getModel(g, True, True, fused=True).forward([in1, in2])
    out = dm.processOutput(g, fused, True, None)
    tie = want["gap"] < NEAR_TIE
    assert ((out["index"] != want["index"]) & ~tie).sum() == 0


@pytest.mark.parametrize("threshold", [0.11, 0.21, 0.0, -0.5])
@pytest.mark.parametrize("n", [5, 33, 289, 1089])
def test_extract_output_standalone_bit_exact(dm, oracle, threshold, n):
    rng = np.random.default_rng(n)
    inp = rng.random((13, 17, n), dtype=np.float32) * (0.5 if n < 64 else 0.24)
    inp[0, 0] = 0
    inp[1, 1] = inp[1, 1, 0]
    r0 = rng.integers(-5, 5, (13, 17)).astype(np.int64)
    s0 = rng.random((13, 17)).astype(np.float32)
    wr, ws, written = oracle.extract_output(inp, threshold, r0, s0)
    ret, sc = r0.copy(), s0.copy()
    nun = dm.extractoutput.extractOutput(inp, sc, threshold, ret)
    np.testing.assert_array_equal(ret, wr)
    np.testing.assert_array_equal(sc, ws)
    assert nun == 13 * 17 - written
    wr, wg = oracle.extract_output_marginalized(inp, threshold, 0.5, r0)
    ret, gd = r0.copy(), np.full((13, 17), 9, np.int64)
    dm.extractoutput.extractOutputMarginalized(inp, threshold, 0.5, ret, gd)
    np.testing.assert_array_equal(ret, wr)
    np.testing.assert_array_equal(gd, wg)


def test_radial_matching_vs_oracle(dm, oracle):
    """c4's matcher: polar maps 10 x 384 x 400 cropped by hWin-1 rows, hWin = 15 (smaller here)."""
    rng = np.random.default_rng(9)
    C, Hp, Wp, hWin = 10, 64, 80, 15
    f2 = rng.standard_normal((C, Hp, Wp)).astype(np.float32)
    f1 = np.roll(f2, -4, axis=1)[:, : Hp - hWin + 1] + 0.05 * rng.standard_normal((C, Hp - hWin + 1, Wp)).astype(np.float32)
    vol = oracle.radial_matching(f1, f2, hWin)
    m = dm.nn.SpatialRadialMatching(hWin)
    np.testing.assert_array_equal(m.forward([f1, f2]), vol)
    idx, mn = oracle.argmin_tie(vol, hWin, 0)
    flow, gmin = m.argmin_flow([f1, f2])
    np.testing.assert_array_equal(flow, (idx - 1).reshape(flow.shape).astype(np.float32))
    np.testing.assert_array_equal(gmin, mn.reshape(flow.shape))
    assert (flow == 4).mean() > 0.99


@pytest.mark.parametrize("maxh,maxw,ratios", [(8, 8, [1, 2]), (8, 8, [1, 2, 4]), (16, 16, [1, 2, 4])])
def test_x2yx_multi_both_modes(dm, oracle, maxh, maxw, ratios):
    g = dm.Geometry(maxh=maxh, maxw=maxw, ratios=ratios, multiscale=True)
    L = dm.multiscaleLength(g)
    x = np.arange(1, L + 1, dtype=np.int64).reshape(4, -1)
    ry, rx = dm.x2yxMulti(g, x)
    for i, (a, b) in enumerate(zip(ry.reshape(-1), rx.reshape(-1))):
        rc, oy, ox = oracle.x2yx_multi_number(maxh, maxw, ratios, i + 1)
        assert (a, b) == (oy, ox)
    xb = np.arange(-2, L + 38 + (L % 2), dtype=np.int64).reshape(2, -1)
    wy, wx = oracle.x2yx_multi2_bugcompat(xb, maxh, maxw, ratios, fill=0)
    gy, gx = dm.x2yxMulti2(g, xb, bug_compat=True)
    np.testing.assert_array_equal(gy, wy)
    np.testing.assert_array_equal(gx, wx)


def test_cascade_add_vs_oracle(dm, oracle):
    rng = np.random.default_rng(12)
    for ratios, k in (([1, 2, 4], 8), ([1, 2], 8), ([1, 2, 4], 16), ([1, 3], 12)):
        inp = rng.random((len(ratios), 37, k, k)).astype(np.float32)
        want = oracle.cascade_add(inp, ratios)
        got = dm.nn.CascadingAddTable(ratios).forward([inp[i] for i in range(len(ratios))])
        for i in range(len(ratios)):
            np.testing.assert_array_equal(got[i], want[i])
    with pytest.raises(dm.DepthMatchError):
        dm.nn.CascadingAddTable([1, 2]).forward([inp[0][:, :7, :7], inp[1][:, :7, :7]])


def _multiscale_oracle(oracle, f1s, f2s, maxh, maxw, ratios):
    K = maxh * maxw
    per = []
    for (f1, f2, r) in zip(f1s, f2s, ratios):
        vol = oracle.spatial_matching(f1, f2, maxh, maxw)
        prob = oracle.neg_softmax(vol).reshape(f1.shape[1], f1.shape[2], K)
        per.append(oracle.upsample_nearest_rows(prob, r).reshape(-1, maxh, maxw))
    casc = oracle.cascade_add(np.stack(per), ratios)
    vec = oracle.ring_join(casc, ratios)
    middle = oracle.yx2x_multi(maxh, maxw, ratios, 0, 0)
    idx, _ = oracle.argmax_tie(vec, vec.shape[1], middle)
    gap = oracle.top2_relgap(vec, vec.shape[1])
    fy = np.empty_like(idx)
    fx = np.empty_like(idx)
    for i, v in enumerate(idx):
        _, fy[i], fx[i] = oracle.x2yx_multi_number(maxh, maxw, ratios, int(v))
    return idx, fy, fx, gap


@pytest.mark.parametrize("ratios", [[1, 2], [1, 2, 4]])
def test_multiscale_extract_vs_oracle(dm, oracle, ratios):
    """c3 in small: per-scale prefiltered maps -> matching -> softmax -> cascade -> ring ->
    argmax (middle tie rule) -> decode."""
    rng = np.random.default_rng(14)
    maxh = maxw = 8
    C, H, W = 10, 32, 48
    f1s, f2s = [], []
    for r in ratios:
        h, w = H // r, W // r
        f2 = rng.standard_normal((C, h + maxh - 1, w + maxw - 1)).astype(np.float32)
        f1 = f2[:, 3:3 + h, 3:3 + w] + 0.4 * rng.standard_normal((C, h, w)).astype(np.float32)
        f1s.append(np.ascontiguousarray(f1))
        f2s.append(f2)
    g = dm.Geometry(maxh=maxh, maxw=maxw, ratios=ratios, multiscale=True, hImg=H, wImg=W,
                    output_extraction_method="max")
    out = dm.getModelMultiscale(g, True, True).forward(list(zip(f1s, f2s)))
    idx, fy, fx, gap = _multiscale_oracle(oracle, f1s, f2s, maxh, maxw, ratios)
    tie = gap < NEAR_TIE * 10
    bad = (out["index"].reshape(-1) != idx) & ~tie
    assert bad.sum() == 0, (bad.sum(), tie.sum())
    ok = ~tie
    np.testing.assert_array_equal(out["flow_y"].reshape(-1)[ok], fy[ok])
    np.testing.assert_array_equal(out["flow_x"].reshape(-1)[ok], fx[ok])
    po = dm.processOutput(g, out, True, None)
    assert po["full"].shape == (2, H, W)


def test_downsample_avg(dm, oracle):
    rng = np.random.default_rng(15)
    img = rng.random((3, 24, 36)).astype(np.float32)
    from depthmatch import api
    c = dm.default_context()
    for r in (1, 2, 4):
        out = np.empty((3, 24 // r, 36 // r), np.float32)
        api.check(c._lib.dm_downsample_avg(c.handle, img.ctypes.data, 3, 24, 36, r, out.ctypes.data))
        np.testing.assert_array_equal(out, oracle.downsample_avg(img, r))


def test_polar_remap_vs_oracle(dm, oracle):
    """c4's remap at reduced size, epipole from radial/gopro.cal scaled."""
    rng = np.random.default_rng(16)
    hImg, wImg, hIn, wIn, wK = 90, 160, 100, 100, 17
    e2 = (641.4552 * wImg / 1280.0, 344.950836 * wImg / 1280.0)
    img = rng.random((3, hImg, wImg)).astype(np.float32)
    rmax = dm.getRMax(hImg, wImg, e2)
    assert rmax == oracle.get_rmax(hImg, wImg, *e2)
    lp, rp = (wK - 1) // 2, math.ceil((wK - 1) / 2)
    omask = oracle.c2p_mask(wIn, hIn, e2[0], e2[1], lp, rp, rmax, 1.0)
    gmask = dm.getC2PMask(wImg, hImg, wIn, hIn, e2[0], e2[1], lp, rp, rmax, 1.0)
    # CUDA's double sin/cos/pow are within 1-2 ulp of glibc's: coordinates agree to float rounding
    np.testing.assert_allclose(gmask, omask, rtol=0, atol=2e-5)
    want = oracle.warp_bilinear(img, omask)
    np.testing.assert_array_equal(dm.cartesian2polar(img, omask), want)          # LUT path, bit-exact
    got = dm.cartesian2polar(img, wdst=wIn, hdst=hIn, xcenter=e2[0], ycenter=e2[1], lpadding=lp,
                             rpadding=rp, rmax=rmax, alpha=1.0)                  # analytic path
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-4)
    # inverse
    pol = want[:, :, lp:lp + wIn]
    om2 = oracle.p2c_mask(wIn, hIn, wImg, hImg, e2[0], e2[1], rmax, 1.0)
    gm2 = dm.getP2CMask(wIn, hIn, wImg, hImg, e2[0], e2[1], rmax, 1.0)
    np.testing.assert_allclose(gm2, om2, rtol=0, atol=5e-5)
    back = oracle.warp_bilinear(pol, om2)
    np.testing.assert_allclose(dm.polar2cartesian(pol, wImg, hImg, e2[0], e2[1], rmax), back, rtol=0,
                               atol=2e-4)
    flow = rng.random((hImg, wImg)).astype(np.float32) * 3
    netp = dict(hImg=hImg, wImg=wImg)
    infty = dm.getRMax(hImg, wImg, e2) * 0.65
    d, c = oracle.flow2depth(flow, e2[0], e2[1], infty)
    gd, gc = dm.flow2depth(netp, flow, e2, 0.65)
    np.testing.assert_array_equal(gc, c)
    np.testing.assert_allclose(gd, d / np.float32(infty), rtol=1e-6)


def test_argument_errors_are_reported_not_crashed(dm):
    a = np.zeros((2, 8, 8), np.float32)
    b = np.zeros((2, 9, 9), np.float32)
    with pytest.raises(dm.DepthMatchError) as e:
        dm.match_extract(a, b, 5, 5)
    assert e.value.status == -1 and "smaller than" in str(e.value)
    with pytest.raises(dm.DepthMatchError):
        dm.match_extract(a, np.zeros((3, 12, 12), np.float32), 5, 5)
    with pytest.raises(dm.DepthMatchError):
        dm.match_extract(a, np.zeros((2, 12, 12), np.float32), 5, 5, prob_threshold=0.05)


def test_north_config_full_size_properties(dm):
    """640x360, 33x33, C=10 (BASELINE north): too big for the oracle in seconds, so check
    size-independent properties: the planted flow is recovered, probabilities are in (0,1],
    soft means lie inside the window, and the exact and FMA paths agree on every index that is
    not a near tie."""
    maxh = maxw = 33
    in1, in2, flow = make_pair(10, 360, 640, maxh, maxw, seed=1234, noise=0.05)
    got = dm.match_extract(in1, in2, maxh, maxw, canvas=(360, 640),
                           want=("index", "min_ssd", "pmax", "soft_yx", "index_thr", "score_thr"))
    cy = 17
    fy = (got["index"] - 1) // maxw + 1 - cy
    fx = (got["index"] - 1) % maxw + 1 - cy
    np.testing.assert_array_equal(fy, flow[0])
    np.testing.assert_array_equal(fx, flow[1])
    assert (got["pmax"] > 0).all() and (got["pmax"] <= 1.0 + 1e-6).all()
    assert (got["soft_yx"] >= 1 - 1e-4).all() and (got["soft_yx"] <= maxh + 1e-4).all()
    np.testing.assert_array_equal(got["flow_full"][0, 16:16 + 328, 16:16 + 608], flow[0])
    # with sigma 0.05 noise the winner takes nearly all the mass: thresholded == WTA
    assert (got["index_thr"] == got["index"]).mean() > 0.999
    ex = dm.match_extract(in1, in2, maxh, maxw, exact=True, want=("index", "min_ssd"))
    np.testing.assert_array_equal(ex["index"], got["index"])
    np.testing.assert_allclose(ex["min_ssd"], got["min_ssd"], rtol=RTOL)


def test_pipelined_host_batch_matches_single_calls(dm):
    """Host batches of >= 4 pairs are cut in chunks over two streams; same results as pair by pair."""
    maxh, maxw, C, N = 9, 17, 10, 7
    f1 = np.empty((N, C, 40, 72), np.float32)
    f2 = np.empty((N, C, 40, 72), np.float32)
    oy, ox = math.ceil(maxh / 2) - 1, math.ceil(maxw / 2) - 1
    H1, W1 = 40 - maxh + 1, 72 - maxw + 1
    for n in range(N):
        in1, in2, _ = make_pair(C, 40, 72, maxh, maxw, seed=50 + n, noise=0.5)
        f1[n] = 0
        f1[n, :, oy:oy + H1, ox:ox + W1] = in1
        f2[n] = in2
    view = f1[:, :, oy:oy + H1, ox:ox + W1]
    want = ("index", "pmax", "index_thr", "score_thr", "soft_yx", "min_ssd", "n_untouched")
    batch = dm.match_extract(view, f2, maxh, maxw, canvas=(40, 72), want=want)
    for n in range(N):
        one = dm.match_extract(view[n], f2[n], maxh, maxw, canvas=(40, 72), want=want)
        for k in one:
            np.testing.assert_array_equal(batch[k][n], one[k], err_msg=k)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configurations at FULL size.  The oracle needs seconds to minutes per pair at
# these sizes, so these check size-independent properties (planted integer flow recovered,
# band partition == whole frame, zero-flow fixed point, idempotence) instead of the volume.
# ---------------------------------------------------------------------------------------------
def test_c2_batch64_320x180_planted_flow(dm):
    """configs[1]: synthetic 320x180 frame pairs, batch 64, single-scale 33x33 window."""
    import torch
    maxh = maxw = 33
    N = 64
    base = [make_pair(10, 180, 320, maxh, maxw, seed=100 + n, noise=0.05, flow_seed=200 + n) for n in range(4)]
    in1 = torch.from_numpy(np.stack([base[n % 4][0] for n in range(N)])).cuda()
    in2 = torch.from_numpy(np.stack([base[n % 4][1] for n in range(N)])).cuda()
    got = dm.match_extract(in1, in2, maxh, maxw, want=("index", "pmax", "score_thr"))
    torch.cuda.synchronize()
    idx = got["index"].cpu().numpy()
    for n in range(N):
        flow = base[n % 4][2]
        np.testing.assert_array_equal((idx[n] - 1) // maxw - 16, flow[0])
        np.testing.assert_array_equal((idx[n] - 1) % maxw - 16, flow[1])
    # identical pairs in the batch give identical results (no cross-pair state)
    for k in ("index", "pmax", "score_thr"):
        v = got[k].cpu().numpy()
        np.testing.assert_array_equal(v[0], v[4])
        np.testing.assert_array_equal(v[3], v[63])


def test_c5_1080p_65x65_row_bands_equal_whole_frame(dm):
    """configs[4]: 1920x1080, 65x65 window, 8 row bands with a 64-row halo: each band
    computed on its own (what a rank does) must reproduce the whole-frame result exactly."""
    import torch
    from depthmatch import parallel
    maxh = maxw = 65
    C, H, W = 10, 1080, 1920
    g = torch.Generator(device="cuda").manual_seed(5)
    in2 = torch.randn((C, H, W), device="cuda", generator=g)
    H1, W1 = H - maxh + 1, W - maxw + 1
    fy, fx = 7, -11   # planted constant flow
    in1 = in2[:, 32 + fy:32 + fy + H1, 32 + fx:32 + fx + W1] + 0.05 * torch.randn((C, H1, W1), device="cuda", generator=g)
    whole = dm.match_extract(in1, in2, maxh, maxw, want=("index", "pmax", "score_thr"))
    idx = whole["index"]
    assert bool(((idx - 1) // maxw - 32 == fy).all()) and bool(((idx - 1) % maxw - 32 == fx).all())
    bands = parallel.row_bands(H1, 8, maxh)
    assert bands[0] == (0, 127, 127 + 64) and bands[-1][1] == H1
    for rank in (0, 3, 7):
        a, b = parallel.band_inputs(in1, in2, bands[rank])
        part = dm.match_extract(a, b, maxh, maxw, want=("index", "pmax", "score_thr"))
        y0, y1, _ = bands[rank]
        for k in ("index", "pmax", "score_thr"):
            assert torch.equal(part[k], whole[k][y0:y1]), (rank, k)


def test_c3_multiscale_640x360_properties(dm):
    """configs[2]: 3 scales {1,2,4}, 8x8 windows, 640x360.  Identical frames -> every pixel picks
    the zero-flow middle index 28; a frame shifted by (dy,dx) = (-2, 3) within the fine window is
    decoded back to that flow at full resolution."""
    import torch
    maxh = maxw = 8
    ratios, C, H, W = [1, 2, 4], 10, 360, 640
    g = dm.Geometry(maxh=maxh, maxw=maxw, ratios=ratios, multiscale=True, hImg=H, wImg=W,
                    output_extraction_method="max")
    gen = torch.Generator(device="cuda").manual_seed(9)

    def pyramid(shifts):
        # shifts[i]: (dy, dx) of scale i in that scale's pixels, or None = unrelated frames
        inp = []
        for r, sh in zip(ratios, shifts):
            h, w = H // r, W // r
            f2 = torch.randn((C, h + maxh - 1, w + maxw - 1), device="cuda", generator=gen)
            if sh is None:
                f1 = torch.randn((C, h, w), device="cuda", generator=gen)
            else:
                f1 = f2[:, 3 + sh[0]:3 + sh[0] + h, 3 + sh[1]:3 + sh[1] + w].clone()
            inp.append((f1, f2))
        return inp

    out = dm.getModelMultiscale(g, True, True).forward(pyramid([(0, 0)] * 3))
    torch.cuda.synchronize()
    assert bool((out["index"] == 28).all())
    assert bool((out["flow_y"] == 0).all()) and bool((out["flow_x"] == 0).all())
    # fine scale sees (2,-2), scale 2 the consistent (1,-1), scale 4 nothing: the cascade puts
    # 2 on the fine entry (2,-2) and at most ~1 anywhere else
    out = dm.getModelMultiscale(g, True, True).forward(pyramid([(2, -2), (1, -1), None]))
    torch.cuda.synchronize()
    assert float((out["flow_y"] == 2).float().mean()) > 0.999
    assert float((out["flow_x"] == -2).float().mean()) > 0.999
    # flow (-4, 4) is outside the fine 8x8 window (rows -3..4): it must come out of scale 2's
    # ring as (-2, 2) x ratio 2
    out = dm.getModelMultiscale(g, True, True).forward(pyramid([None, (-2, 2), (-1, 1)]))
    torch.cuda.synchronize()
    assert float((out["flow_y"] == -4).float().mean()) > 0.99
    assert float((out["flow_x"] == 4).float().mean()) > 0.99
    idx = out["index"].cpu().numpy()
    ry, rx = dm.x2yxMulti(g, idx)
    np.testing.assert_array_equal(ry, out["flow_y"].cpu().numpy())
    np.testing.assert_array_equal(rx, out["flow_x"].cpu().numpy())


def test_c4_radial_polar_400x400(dm, oracle):
    """configs[3]: 640x360 frames remapped to 400x400 polar maps around the gopro.cal epipole
    (+16 circular pad columns for a 17-wide kernel), then the 1-D radial search, hWin = 15.
    A polar map shifted by 5 rows must give radial flow 5; polar -> cartesian -> polar of a
    smooth image is close to the identity inside the valid disc."""
    import torch
    hImg, wImg, hIn, wIn, wK, hWin, C = 360, 640, 400, 400, 17, 15, 10
    e2 = (641.4552 * wImg / 1280.0, 344.950836 * wImg / 1280.0)
    rmax = dm.getRMax(hImg, wImg, e2)
    yy, xx = np.meshgrid(np.arange(hImg), np.arange(wImg), indexing="ij")
    img = np.stack([np.sin(xx / 23.0) + np.cos(yy / 17.0), np.cos(xx / 31.0) * np.sin(yy / 13.0),
                    0.002 * xx + 0.003 * yy]).astype(np.float32)
    lp, rp = (wK - 1) // 2, math.ceil((wK - 1) / 2)
    pol = dm.cartesian2polar(img, wdst=wIn, hdst=hIn, xcenter=e2[0], ycenter=e2[1], lpadding=lp,
                             rpadding=rp, rmax=rmax, alpha=1.0)
    assert pol.shape == (3, hIn, wIn + 16)
    np.testing.assert_array_equal(pol[:, :, :lp], pol[:, :, wIn:wIn + lp])   # circular padding
    back = dm.polar2cartesian(pol[:, :, lp:lp + wIn], wImg, hImg, e2[0], e2[1], rmax)
    r = np.hypot(xx - e2[0], yy - e2[1])
    # away from the epipole, the frame border (polar samples beyond it are clamped) and the
    # theta = 0 seam (image.warp does not wrap the last polar column)
    inside = (r > 8) & (xx > 8) & (xx < wImg - 9) & (yy > 8) & (yy < hImg - 9)
    inside &= ~((np.abs(yy - e2[1]) < 6) & (xx > e2[0]))
    err = np.abs(back - img)[:, inside]
    assert err.mean() < 0.01 and np.quantile(err, 0.999) < 0.1
    # matcher on 10-channel polar feature maps: in2 10 x 384 x 400, in1 10 x 370 x 400
    f2 = torch.randn((C, 384, 400), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    f1 = f2[:, 5:5 + 384 - hWin + 1].clone()
    flow, mn = dm.nn.SpatialRadialMatching(hWin).argmin_flow([f1, f2])
    torch.cuda.synchronize()
    assert flow.shape == (370, 400) and bool((flow == 5).all()) and bool((mn == 0).all())
    depth, conf = dm.flow2depth(dict(hImg=370, wImg=400), flow.cpu().numpy(), (200.0, 185.0), 0.65)
    assert conf[185, 200] == 0 and depth.max() <= 1.0 + 1e-6


# ---------------------------------------------------------------------------------------------
# "next" rows: the steps right after the matching path, bit-exact with the oracle
# (which is itself pinned against the reference's inline C, tests/test_oracle_ref.py)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["med", "max"])
@pytest.mark.parametrize("k", [3, 5])
def test_post_process_image_vs_oracle(dm, oracle, method, k):
    rng = np.random.default_rng(k + 7)
    flow = np.clip(np.rint(rng.normal(0, 3, (2, 90, 160))), -7, 8).astype(np.float32)
    if method == "med":
        flow = flow + rng.random((2, 90, 160)).astype(np.float32)
    else:
        flow = flow + (rng.random((2, 90, 160)).astype(np.float32) - 0.5) * 0.8
    mask = (rng.random((90, 160)) > 0.3).astype(np.float32)
    mask[20:30, 40:60] = 0
    want = oracle.post_process_image(flow, mask, k, method)
    np.testing.assert_array_equal(dm.postProcessImage(flow, mask, k, method), want)


def test_enlarge_mask_radial_and_drone_depth_vs_oracle(dm, oracle):
    rng = np.random.default_rng(21)
    for ix, iy in ((16, 16), (3, 5), (1, 1)):
        mask = (rng.random((180, 320)) > 0.35).astype(np.float32)
        mask[7] = 0
        mask[:, 11] = 0
        want = oracle.enlarge_mask(mask, ix, iy)
        got = dm.enlargeMask(mask.copy(), ix, iy)
        np.testing.assert_array_equal(got, want)
    flow = (rng.normal(0, 2, (2, 180, 320)) * (rng.random((2, 180, 320)) > 0.2)).astype(np.float32)
    g = dm.Geometry(wImg=320, hImg=180)
    wr, wc = oracle.radial_depth(flow, 125.8, 155.2, 160.0)
    gr, gc = dm.radial(g, flow, 125.8, 155.2)
    np.testing.assert_array_equal(gr, wr)
    np.testing.assert_array_equal(gc, wc)
    xflow = np.clip(rng.normal(0, 3, (180, 320)), -8, 11).astype(np.float32)
    mask = (rng.random((180, 320)) > 0.2).astype(np.float32)
    wd, wcf = oracle.depth_from_xflow(xflow, mask, 0.37)
    gd, gcf = dm.computeDepthMapFromFlow(xflow, mask, 0.37)
    np.testing.assert_array_equal(gd, wd)
    np.testing.assert_array_equal(gcf, wcf)


# ---------------------------------------------------------------------------------------------
# "next" row 3: the feature extractor (getFilter / getMultiscalePrefilter), fp32 within 1e-4
# ---------------------------------------------------------------------------------------------
def _oracle_layers(flt, dm):
    layers = []
    for m in flt.modules:
        if isinstance(m, dm.nn.Tanh):
            layers[-1]["tanh"] = True
        else:
            layers.append(dict(weight=m.weight, bias=m.bias, conn=m.connTable))
    return layers


def _assert_close_fp32(got, want, rel=1e-4):
    scale = float(np.abs(want).max())
    np.testing.assert_allclose(got, want, rtol=rel, atol=rel * scale)


def test_filter_c1_geometry_vs_oracle(dm, oracle):
    """c1's filter {3,5,5,8} tanh {4,16,16,10}: a full convolution, then a
    SpatialConvolutionMap with nn.tables.random(8, 10, 4)."""
    rng = np.random.default_rng(1)
    g = dm.Geometry(layers=[[3, 5, 5, 8], [4, 16, 16, 10]])
    flt = dm.getFilter(g, rng)
    assert flt.modules[2].connTable.shape == (40, 2) and isinstance(flt.modules[1], dm.nn.Tanh)
    for o in range(10):   # every output plane reads 4 distinct input planes
        rows = flt.modules[2].connTable[flt.modules[2].connTable[:, 1] == o + 1, 0]
        assert len(rows) == 4 and len(set(rows)) == 4
    img = rng.random((3, 90, 160)).astype(np.float32)
    want = oracle.filter_forward(img, _oracle_layers(flt, dm))
    got = flt.forward(img)
    assert got.shape == want.shape == (10, 71, 141)
    _assert_close_fp32(got, want)
    # batch of both frames = two single calls, bit for bit; device tensors too
    img2 = rng.random((3, 90, 160)).astype(np.float32)
    both = flt.forward(np.stack([img, img2]))
    np.testing.assert_array_equal(both[0], got)
    np.testing.assert_array_equal(both[1], flt.forward(img2))
    import torch
    dev = flt.forward(torch.from_numpy(np.stack([img, img2])).cuda())
    assert dev.is_cuda
    np.testing.assert_array_equal(dev.cpu().numpy(), both)


@pytest.mark.parametrize("shape,layers", [
    ((3, 40, 70), [[3, 1, 17, 5], [5, 17, 1, 10]]),               # radial net: 1x17 then 17x1
    ((3, 40, 70), [[3, 3, 3, 4], "tanh", [4, 7, 2, 6], "tanh"]),  # ragged kernels, trailing tanh
    ((1, 33, 65), [[1, 1, 1, 2]]),                                  # 1x1
    ((10, 50, 60), [[10, 17, 17, 10]]),                             # widest footprint used anywhere
])
def test_radial_filter_shapes_vs_oracle(dm, oracle, shape, layers):
    rng = np.random.default_rng(2)
    flt = dm.getRadialFilter(dict(layers=layers), rng)
    img = rng.standard_normal(shape).astype(np.float32)
    want = oracle.filter_forward(img, _oracle_layers(flt, dm))
    _assert_close_fp32(flt.forward(img), want)


def test_multiscale_prefilter_and_raw_frames_end_to_end(dm, oracle):
    """c3 from raw frames: downsample -> zero padding -> shared {3,5,5,10} filter per scale,
    and getModel(prefiltered=false): filter on both patches, then the fused matcher."""
    rng = np.random.default_rng(3)
    g = dm.Geometry(layers=[[3, 5, 5, 10]], ratios=[1, 2, 4], multiscale=True, share_filters=True,
                    wPatch2=5, hPatch2=5, maxh=8, maxw=8)
    flt = dm.getFilter(g, rng)
    pre = dm.getMultiscalePrefilter(g, flt)
    img = rng.random((3, 64, 96)).astype(np.float32)
    feats = pre(img)
    layers = _oracle_layers(flt, dm)
    for r, f in zip(g.ratios, feats):
        small = oracle.downsample_avg(img, r) if r != 1 else img
        want = oracle.filter_forward(small, layers, pads=(2, 2, 2, 2))
        assert f.shape == (10, 64 // r, 96 // r)
        _assert_close_fp32(f, want)
    # single scale, raw frames in, flow out
    g1 = dm.Geometry(layers=[[3, 5, 5, 8], [4, 16, 16, 10]], maxh=9, maxw=9, hImg=80, wImg=120)
    flt1 = dm.getFilter(g1, rng)
    base = rng.random((3, 90, 130)).astype(np.float32)
    fr2 = base[:, 5:85, 5:125]
    fr1 = base[:, 3:83, 7:127]          # frame 1 pixel (y,x) = frame 2 pixel (y-2, x+2)
    model = dm.getModel(g1, True, False, fused=True, filter=flt1)
    out = model.forward(dm.prepareInput(g1, fr1, fr2))
    f1 = oracle.filter_forward(np.ascontiguousarray(dm.prepareInput(g1, fr1, fr2)[0]), _oracle_layers(flt1, dm))
    f2 = oracle.filter_forward(fr2, _oracle_layers(flt1, dm))
    prob = oracle.neg_softmax(oracle.spatial_matching(f1, f2, 9, 9))
    idx, _ = oracle.argmax_tie(prob, 81, dm.getMiddleIndex(g1))
    gap = oracle.top2_relgap(prob, 81)
    got = np.asarray(out["index"]).reshape(-1)
    bad = (got != idx) & (gap >= 1e-3)   # features differ by ~1e-6 relative: allow near-ties
    assert bad.sum() == 0
    dy, dx = dm.x2yx(g1, got.reshape(f1.shape[1:]))
    assert np.median(dy) - 5 == -2 and np.median(dx) - 5 == 2


def test_saved_model_loads_and_runs(dm, oracle, tmp_path):
    """next row 4: saveModel -> loadModel (version-9 Torch7 table) -> forward gives the flow of
    the model that was saved (weights travel; the connection table is re-drawn on load, as in the
    reference, so the second layer is a full convolution here to make the outputs comparable)."""
    rng = np.random.default_rng(8)
    g = dm.Geometry(layers=[[3, 5, 5, 6], [6, 7, 7, 10]], maxh=9, maxw=9, maxhHR=9, maxwHR=9, maxhGT=9, maxwGT=9,
                    hImg=60, wImg=90, output_extraction_method="max")
    learning = dict(rate=0.01, rate_decay=0, weight_decay=0, first_image=0, delta=1, num_images=2)
    model = dm.getModel(g, True, False, fused=True, rng=rng)
    path = dm.saveModel(str(tmp_path), "m", g, learning, model, 1)
    loaded = dm.loadModel(path, True, False, fused=True, rng=np.random.default_rng(99))
    # biases are not part of a version-9 file (opticalflow_model_io.lua:151): copy them by hand
    for a, b in zip(loaded["model"].filter.modules, model.filter.modules):
        if hasattr(a, "bias"):
            assert np.array_equal(a.weight, b.weight) and not np.array_equal(a.bias, b.bias)
            a.bias[...] = b.bias
    loaded["model"].filter.reset_weights()
    fr = rng.random((2, 3, 60, 90)).astype(np.float32)
    want = model.forward(dm.prepareInput(g, fr[0], fr[1]))
    got = loaded["model"].forward(dm.prepareInput(loaded["geometry"], fr[0], fr[1]))
    np.testing.assert_array_equal(got["index"], want["index"])
    np.testing.assert_array_equal(got["pmax"], want["pmax"])


def test_async_host_call_and_cropped_view_equal_the_synchronous_call(dm):
    """DM_FLAG_ASYNC (results valid after ctx.synchronize()) and the packed 3-D copy of a cropped
    frame-1 view give exactly what the plain call gives."""
    import torch
    rng = np.random.default_rng(17)
    N, C, H, W, mh = 8, 4, 40, 56, 9
    f2 = torch.from_numpy(rng.standard_normal((N, C, H, W)).astype(np.float32)).pin_memory().numpy()
    f1 = torch.from_numpy(rng.standard_normal((N, C, H, W)).astype(np.float32)).pin_memory().numpy()
    in1 = f1[:, :, 4:4 + H - mh + 1, 4:4 + W - mh + 1]          # prepareInput's narrow: a strided view
    want = ("index", "pmax", "score_thr")
    ref = dm.match_extract(np.ascontiguousarray(in1), f2, mh, mh, want=want)
    ctx = dm.Context(0)
    out = {k: torch.empty(tuple(v.shape), dtype=torch.from_numpy(v).dtype).pin_memory().numpy()
           for k, v in ref.items()}
    for k in out:
        out[k][...] = 0
    res = dm.match_extract(in1, f2, mh, mh, want=want, ctx=ctx, out=out, async_=True)
    ctx.synchronize()
    for k in want:
        np.testing.assert_array_equal(res[k], ref[k])
    with pytest.raises(dm.DepthMatchError):      # a conversion copy would not outlive the call
        dm.match_extract(in1.astype(np.float64), f2, mh, mh, want=want, ctx=ctx, out=out, async_=True)
    assert ctx.launch_count() > 0


def test_dot_and_difference_forms_of_the_ssd_agree(dm, oracle, monkeypatch):
    """Large calls form the SSD as |a|^2+|b|^2-2a.b when the norms allow it (DM_FLAG_DIFF_SSD
    forces sum (a-b)^2).  Both must satisfy the parity bars against the oracle; min_ssd is
    re-scored in the difference form; large norms fall back on the device.  the ssd_form option
    forces the dot kernel on inputs small enough for the oracle."""
    dm.default_context().set_option("ssd_form", "dot")
    maxh, maxw = 9, 11
    in1, in2, _ = make_pair(10, 60, 90, maxh, maxw, seed=31, noise=0.3)
    K = maxh * maxw
    want = ("index", "min_ssd", "pmax", "index_thr", "score_thr", "soft_yx")
    dot = dm.match_extract(in1, in2, maxh, maxw, want=want)
    dif = dm.match_extract(in1, in2, maxh, maxw, want=want, diff_form=True)
    vol = oracle.spatial_matching(in1, in2, maxh, maxw)
    prob = oracle.neg_softmax(vol)
    idx, pmax = oracle.argmax_tie(prob, K, dm.getMiddleIndex(dm.Geometry(maxh=maxh, maxw=maxw)))
    gap = oracle.top2_relgap(prob, K)
    for got in (dot, dif):
        bad = (got["index"].reshape(-1) != idx) & (gap >= 1e-5)
        assert bad.sum() == 0
        np.testing.assert_allclose(got["pmax"].reshape(-1), pmax, rtol=1e-4)
        np.testing.assert_allclose(got["min_ssd"].reshape(-1), vol.reshape(-1, K).min(-1), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(dot["soft_yx"], dif["soft_yx"], rtol=1e-4)
    same = dot["index"] == dif["index"]
    assert same.mean() > 0.999
    # an exact copy: the dot form would read ~1e-6 where the SSD is exactly 0
    f2 = in2.copy()
    f1 = np.ascontiguousarray(f2[:, 3:3 + in1.shape[1], 5:5 + in1.shape[2]])
    got = dm.match_extract(f1, f2, maxh, maxw, want=("index", "min_ssd"))
    assert (got["min_ssd"] == 0).all() and (got["index"] == 3 * maxw + 5 + 1).all()
    # flat frames: every displacement ties bit for bit, the zero-flow rule picks the middle
    flat1 = np.full((10, 40, 50), 0.7, np.float32)
    flat2 = np.full((10, 40 + maxh - 1, 50 + maxw - 1), 0.7, np.float32)
    got = dm.match_extract(flat1, flat2, maxh, maxw, want=("index", "pmax"))
    assert (got["index"] == dm.getMiddleIndex(dm.Geometry(maxh=maxh, maxw=maxw))).all()
    np.testing.assert_allclose(got["pmax"], 1.0 / K, rtol=1e-5)
    # automatic mode on a call large enough for the dot form: unit-variance features take it
    # (results differ from the difference form in the last bits), norms beyond the bound make the
    # device-side switch run the difference kernel, bit for bit
    dm.default_context().set_option("ssd_form", "auto")
    in1, in2, _ = make_pair(10, 260, 260, 33, 33, seed=32, noise=0.2)
    a = dm.match_extract(in1, in2, 33, 33, want=("index", "pmax"))
    b = dm.match_extract(in1, in2, 33, 33, want=("index", "pmax"), diff_form=True)
    assert (a["index"] == b["index"]).mean() > 0.999 and not np.array_equal(a["pmax"], b["pmax"])
    np.testing.assert_allclose(a["pmax"], b["pmax"], rtol=1e-4)
    big1, big2 = in1 * 30, in2 * 30
    a = dm.match_extract(big1, big2, 33, 33, want=("index", "min_ssd", "pmax"))
    b = dm.match_extract(big1, big2, 33, 33, want=("index", "min_ssd", "pmax"), diff_form=True)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])


def test_depth_estimation_api_next_frame_depth(dm, oracle):
    """depth_estimation_api.lua's nextFrameDepth from the scaled frame on: ego-motion warp of the
    previous features, filter, prepareInput, matching, soft-max, 'mean' extraction, enlargeMask,
    mask composition -- against the same chain built from the oracle's pieces."""
    rng = np.random.default_rng(41)
    hImg, wImg, maxh, maxw = 60, 84, 7, 9
    g = dm.Geometry(layers=[[3, 5, 5, 6]], maxh=maxh, maxw=maxw, hImg=hImg, wImg=wImg)
    flt = dm.getFilter(g, rng)
    layers = _oracle_layers(flt, dm)
    base = rng.random((3, hImg + 8, wImg + 8)).astype(np.float32)
    frames = [np.ascontiguousarray(base[:, 4 + s:4 + s + hImg, 4 - s:4 - s + wImg]) for s in (0, 1, 2)]
    K = np.array([[70.0, 0, wImg], [0, 70.0, hImg], [0, 0, 1]])       # full-size camera; Khalf is used
    a = 0.01
    R = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])
    api = dm.DepthEstimationAPI(g, flt, K=K, first_frame=frames[0])
    Khalf = K * 0.5
    Khalf[2, 2] = 1
    Hm = Khalf @ R @ np.linalg.inv(Khalf)
    last = oracle.filter_forward(frames[0], layers)
    for t in (1, 2):
        im, xflow, mask = api.nextFrameDepth(frames[t], R=R, nFound=100, nInliers=90)
        # the oracle's chain
        warped, wmask = oracle.warp_homography(last, Hm)
        filt = oracle.filter_forward(frames[t], layers)
        in1, in2 = dm.prepareInput(g, warped, filt)
        prob = oracle.neg_softmax(oracle.spatial_matching(np.ascontiguousarray(in1), in2, maxh, maxw))
        ym, xm = oracle.soft_mean(prob, maxh, maxw)
        h1, w1 = in1.shape[1:]
        pm = oracle.marginal_x(prob, maxh, maxw).reshape(h1, w1, maxh)
        _, sc, _ = oracle.extract_output(pm, 0.11, np.zeros((h1, w1), np.int64), np.zeros((h1, w1), np.float32))
        conf = np.zeros((hImg, wImg), np.float32)
        hoff, woff = (hImg - h1) // 2, (wImg - w1) // 2
        conf[hoff:hoff + h1, woff:woff + w1] = sc > 0
        want_x = np.zeros((hImg, wImg), np.float32)
        want_x[hoff:hoff + h1, woff:woff + w1] = xm.reshape(h1, w1) - math.ceil(maxw / 2)
        m = oracle.enlarge_mask(wmask, math.ceil((wImg - w1) / 2), math.ceil((hImg - h1) / 2))
        mh, mw = m.shape
        want_mask = np.zeros((hImg, wImg), np.float32)
        oy, ox = (hImg - mh) // 2 - 1, (wImg - mw) // 2 - 1
        want_mask[oy:oy + mh, ox:ox + mw] = m
        want_mask *= conf
        assert xflow.shape == (hImg, wImg) and mask.shape == (hImg, wImg)
        np.testing.assert_allclose(xflow, want_x, rtol=0, atol=2e-3)
        assert (mask != want_mask).mean() < 2e-3
        assert 0.05 < mask.mean() < 1.0
        last = filt
    # a bad frame (few inliers): zero flow, zero mask, the state still advances
    _, xflow, mask = api.nextFrameDepth(frames[0], R=R, nFound=100, nInliers=5)
    assert not xflow.any() and not mask.any()
    # the homography warp alone, against the oracle (identity and a rotation)
    src = rng.random((4, 30, 44)).astype(np.float32)
    for Hq in (np.eye(3), Hm, np.array([[1, 0, 2.5], [0, 1, -1.25], [0, 0, 1.0]])):
        want, wm = oracle.warp_homography(src, Hq)
        got, gm = dm.warpHomography(src, Hq)
        np.testing.assert_array_equal(gm, wm)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)


def test_radial_pipeline_from_frames_vs_oracle_chain(dm, oracle):
    """Config 4 in small, from RGB frames: polar remap (LUT and analytic), shared 1x17 / 17x1
    filter, radial matcher + argmin, back to cartesian, flow2depth -- against the oracle's chain.
    The epipole comes from the reference's gopro.cal (tests/golden/ref_calibration.npz)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_calibration.npz"))
    cal = dm.torch7io.loads(z["gopro"].tobytes())
    netp = dict(wImg=160, hImg=90, wInput=96, hInput=100, wKernel=17, hKernel=17, hWin=9,
                layers=[[3, 1, 17, 5], [5, 17, 1, 10]])
    e2 = (float(cal.K[0, 2]) * netp["wImg"] / cal.wImg, float(cal.K[1, 2]) * netp["wImg"] / cal.wImg)
    rng = np.random.default_rng(51)
    flt = dm.getRadialFilter(netp, rng)
    layers = _oracle_layers(flt, dm)
    base = rng.random((3, 90, 160)).astype(np.float32)
    prev = base
    img = np.clip(base + 0.02 * rng.standard_normal(base.shape), 0, 1).astype(np.float32)
    got = dm.RadialTester(netp, flt, use_masks=True).forward(prev, img, e2)
    # oracle chain (radial/test_radial_opticalflow.lua:183-221)
    rmax = oracle.get_rmax(90, 160, *e2)
    m = oracle.c2p_mask(96, 100, e2[0], e2[1], 8, 8, rmax, 1.0)
    p_img, p_prev = oracle.warp_bilinear(img, m), oracle.warp_bilinear(prev, m)
    f_prev = oracle.filter_forward(np.ascontiguousarray(p_prev[:, :100 - 9 + 1]), layers)
    f_img = oracle.filter_forward(p_img, layers)
    vol = oracle.radial_matching(f_prev, f_img, 9)
    idx, _ = oracle.argmin_tie(vol, 9, 0)
    gap = np.sort(vol.reshape(-1, 9), -1)
    safe = ((gap[:, 1] - gap[:, 0]) > 1e-4 * np.maximum(gap[:, 0], 1e-6)).reshape(f_prev.shape[1:])
    want_idx = (idx - 1).reshape(f_prev.shape[1:]).astype(np.float32)
    assert got["polar_flow"].shape == want_idx.shape == (76, 96)
    assert (got["polar_flow"] != want_idx)[safe].sum() == 0 and safe.mean() > 0.9
    hPolar = 100 - 17 - 9 + 2
    k = hPolar / 100
    m2 = oracle.p2c_mask(96, hPolar, int(160 * k), int(90 * k), e2[0] * k, e2[1] * k, rmax * k, 1.0)
    cart = oracle.warp_bilinear(got["polar_flow"][None], m2)[0]
    np.testing.assert_allclose(got["cart_flow"], cart, rtol=0, atol=1e-4)
    kout = dm.getKOutput(netp)
    infty = rmax * 0.65
    d, c = oracle.flow2depth(got["cart_flow"], e2[0] * kout, e2[1] * kout, infty)
    np.testing.assert_array_equal(got["confs"], c)
    np.testing.assert_allclose(got["depth"], d / np.float32(infty), rtol=1e-6)
    # the analytic remaps (no LUT in HBM) agree with the LUT path up to the libm of the device
    ana = dm.RadialTester(netp, flt, use_masks=False).forward(prev, img, e2)
    assert (ana["polar_flow"] != got["polar_flow"])[safe].mean() < 0.01


def test_multiscale_from_raw_frames(dm, oracle):
    """Config 3 from frames: multiscaleInputs (average, padding, shared filter per scale) feeds
    getModelMultiscale; the result equals the oracle's cascade on the same feature maps and a
    planted shift of the second frame is recovered away from the zero-padded border."""
    rng = np.random.default_rng(61)
    g = dm.Geometry(maxh=8, maxw=8, ratios=[1, 2, 4], multiscale=True, hImg=64, wImg=96, wPatch2=5, hPatch2=5,
                    layers=[[3, 5, 5, 10]], share_filters=True, output_extraction_method="max")
    flt = dm.getFilter(g, rng)
    base = rng.random((3, 64 + 16, 96 + 16)).astype(np.float32)
    img2 = np.ascontiguousarray(base[:, 8:72, 8:104])
    # frame-2 content at (+4, +4): an integer displacement at every scale (with random filters the
    # un-normalised cascade only recovers flows all the scales agree on)
    img1 = np.ascontiguousarray(base[:, 8 + 4:72 + 4, 8 + 4:104 + 4])
    inp = dm.multiscaleInputs(g, flt, img1, img2)
    for (f1, f2), r in zip(inp, g.ratios):
        assert f1.shape == (10, 64 // r, 96 // r) and f2.shape == (10, 64 // r + 7, 96 // r + 7)
    out = dm.getModelMultiscale(g, True, True).forward(inp)
    # oracle on the same maps
    f1s, f2s = [np.asarray(a) for a, _ in inp], [np.asarray(b) for _, b in inp]
    idx, fy, fx, gap = _multiscale_oracle(oracle, f1s, f2s, 8, 8, g.ratios)
    tie = gap < NEAR_TIE * 10
    assert ((np.asarray(out["index"]).reshape(-1) != idx) & ~tie).sum() == 0 and tie.mean() < 0.05
    inner = (slice(20, 44), slice(24, 72))
    assert np.median(np.asarray(out["flow_y"])[inner]) == 4 and np.median(np.asarray(out["flow_x"])[inner]) == 4
    # the same through getModelMultiscale(prefiltered=false)
    out2 = dm.getModelMultiscale(g, True, False, filter=flt).forward([img1, img2])
    np.testing.assert_array_equal(np.asarray(out2["index"]), np.asarray(out["index"]))


def test_winner_take_all_only_path_equals_the_full_path(dm, oracle, monkeypatch):
    """When no probability is asked for (index / flow / min_ssd only) the kernel skips the
    soft-max: same indices, same canvas, same minima as the full epilogue, in every SSD form."""
    maxh, maxw = 7, 9
    in1, in2, _ = make_pair(10, 50, 70, maxh, maxw, seed=71, noise=0.4)
    in1[:, 10:20, 10:30] = 0.25          # a flat patch against a flat patch: ties -> zero flow
    in2[:, 10:26, 10:38] = 0.25
    for form in ("diff", "dot"):
        dm.default_context().set_option("ssd_form", form)
        full = dm.match_extract(in1, in2, maxh, maxw, canvas=(50, 70), want=("index", "min_ssd", "pmax"))
        wta = dm.match_extract(in1, in2, maxh, maxw, canvas=(50, 70), want=("index", "min_ssd"))
        np.testing.assert_array_equal(wta["index"], full["index"])
        np.testing.assert_array_equal(wta["min_ssd"], full["min_ssd"])
        np.testing.assert_array_equal(wta["flow_full"], full["flow_full"])
        assert (wta["index"][12:18, 12:28] == dm.getMiddleIndex(dm.Geometry(maxh=maxh, maxw=maxw))).all()
    dm.default_context().set_option("ssd_form", "auto")
    vol = oracle.spatial_matching(in1, in2, maxh, maxw).reshape(-1, maxh * maxw)
    prob = oracle.neg_softmax(vol)
    idx, _ = oracle.argmax_tie(prob, maxh * maxw, dm.getMiddleIndex(dm.Geometry(maxh=maxh, maxw=maxw)))
    gap = oracle.top2_relgap(prob, maxh * maxw)
    got = dm.match_extract(in1, in2, maxh, maxw, want=("index",), exact=True)["index"].reshape(-1)
    assert ((got != idx) & (gap >= 1e-5)).sum() == 0


def test_randomised_shapes_forms_and_outputs(dm):
    """tests/fuzz_parity.py: 30 random (channels, window, frame size, SSD form, noise, flat
    regions) cases of the fused path -- every output -- against the oracle."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "fuzz_parity.py"), "5", "30"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "failures: 0" in out.stdout


def test_randomised_volume_and_filter_cases(dm):
    """tests/fuzz_volume.py: random shapes of the volume path (SSD and soft-max, exact and FMA)
    and of single filter layers (full and connection-table, padding, tanh) against the oracle."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "fuzz_volume.py"), "7", "24"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "failures: 0" in out.stdout


def test_feature_stream_equals_independent_pairs(dm):
    """FeatureStream (every frame uploaded once, the previous one resident) returns, pair by
    pair, exactly what the independent-pair call returns."""
    import torch
    rng = np.random.default_rng(81)
    C, H, W, mh, mw, B = 4, 40, 56, 5, 7, 3
    frames = rng.standard_normal((1 + 3 * B, C, H, W)).astype(np.float32)
    s = dm.FeatureStream(mh, mw, C, H, W, batch=B, want=("index", "pmax", "score_thr"))
    s.prime(frames[0])
    pinned = [torch.from_numpy(frames[1 + k * B:1 + (k + 1) * B]).pin_memory() for k in range(3)]
    handles = [s.push(p) for p in pinned[:2]]
    got = [{k: v.copy() for k, v in h.wait().items()} for h in handles]     # buffers are reused later
    got.append({k: v.copy() for k, v in s.push(pinned[2]).wait().items()})
    s.synchronize()
    oy, ox = 2, 3
    for k in range(3):
        for i in range(B):
            t = k * B + i                      # pair (frame t, frame t + 1)
            in1 = np.ascontiguousarray(frames[t][:, oy:oy + H - mh + 1, ox:ox + W - mw + 1])
            ref = dm.match_extract(in1, frames[t + 1], mh, mw, canvas=(H, W), want=("index", "pmax", "score_thr"))
            for name in ("index", "pmax", "score_thr", "flow_full"):
                np.testing.assert_array_equal(got[k][name][i], ref[name], err_msg="%s batch %d pair %d" % (name, k, i))


def test_randomised_multiscale_extract_postprocess_and_warps(dm):
    """tests/fuzz_misc.py: random geometries of the multiscale model, stand-alone extractOutput,
    postProcessImage / enlargeMask and the LUT / homography warps against the oracle."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "fuzz_misc.py"), "9", "20"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "failures: 0" in out.stdout


def test_device_calls_are_stream_ordered_and_graph_capturable(dm):
    """A call on device buffers does no host synchronisation, allocation or pageable copy once
    the arena is warm: it can be captured in a CUDA graph and replayed with the same result."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(1)
    in2 = torch.randn((2, 10, 60, 140), device="cuda", generator=g)
    in1 = (in2[:, :, 4:4 + 52, 4:4 + 132] + 0.05 * torch.randn((2, 10, 52, 132), device="cuda", generator=g)).contiguous()
    ctx = dm.Context(0)
    want = ("index", "pmax", "score_thr")
    out = {"index": torch.empty((2, 52, 132), dtype=torch.int64, device="cuda"),
           "pmax": torch.empty((2, 52, 132), device="cuda"), "score_thr": torch.empty((2, 52, 132), device="cuda")}
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            dm.match_extract(in1, in2, 9, 9, want=want, ctx=ctx, out=out)
    s.synchronize()
    ref = {k: v.clone() for k, v in out.items()}
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        dm.match_extract(in1, in2, 9, 9, want=want, ctx=ctx, out=out)
    for v in out.values():
        v.zero_()
    graph.replay()
    torch.cuda.synchronize()
    for k in out:
        assert torch.equal(out[k], ref[k]), k


def test_fused_mean_extraction_with_marginal_confidence(dm, oracle):
    """'mean' extraction (OutputExtractor + getOutputConfidences2's marginal-over-x threshold test,
    opticalflow_model.lua:171-199) out of the fused kernel, against the oracle on the volume."""
    maxh, maxw = 15, 5
    in1, in2, _ = make_pair(10, 56, 150, maxh, maxw, seed=91, noise=0.3)
    # left half: low-contrast features -> flat soft-max, every row marginal ~1/15 < 0.11
    in1[:, :, :70] *= 0.05
    in2[:, :, :72] *= 0.05
    got = dm.match_extract(in1, in2, maxh, maxw, want=("soft_yx", "conf_marginal", "pmax"))
    prob = oracle.neg_softmax(oracle.spatial_matching(in1, in2, maxh, maxw))
    h1, w1 = in1.shape[1:]
    ym, xm = oracle.soft_mean(prob, maxh, maxw)
    np.testing.assert_allclose(got["soft_yx"][0].reshape(-1), ym, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got["soft_yx"][1].reshape(-1), xm, rtol=1e-4, atol=1e-4)
    pm = oracle.marginal_x(prob, maxh, maxw).reshape(h1, w1, maxh)
    _, sc, _ = oracle.extract_output(pm, 0.11, np.zeros((h1, w1), np.int64), np.zeros((h1, w1), np.float32))
    near = np.abs(pm.max(-1) - 0.11) < 1e-4
    want = (sc > 0).astype(np.float32)
    assert 0.02 < want.mean() < 0.98           # both classes are exercised
    np.testing.assert_array_equal(got["conf_marginal"][~near], want[~near])
    with pytest.raises(dm.DepthMatchError):
        dm.match_extract(in1, in2, maxh, maxw, want=("index", "conf_marginal"))
    # the processOutput mirror picks it up for output_extraction_method = 'mean'
    g = dm.Geometry(maxh=maxh, maxw=maxw, hImg=56, wImg=150, output_extraction_method="mean")
    po = dm.processOutput(g, dm.getModel(g, True, True, fused=True).forward([in1, in2]), True, None)
    np.testing.assert_array_equal(np.asarray(po["confidences"])[~near], want[~near] > 0)


def test_no_writes_outside_the_output_buffers(dm):
    """Guard bands: every output of the fused call and the volume are carved out of sentinel-filled
    device buffers; after the calls the bands are intact (compute-sanitizer is not available on
    this pool, so out-of-bounds stores are looked for this way), for ragged shapes and both tile
    configurations."""
    import ctypes as C
    import torch
    from depthmatch import _lib
    rng = np.random.default_rng(101)
    G = 4096   # guard band, elements
    for (n, c, h2, w2, mh, mw) in [(1, 10, 23, 141, 5, 9), (3, 4, 40, 67, 7, 1), (2, 16, 19, 300, 3, 17),
                                   (1, 3, 9, 5, 2, 1), (5, 10, 52, 135, 9, 9)]:
        h1, w1 = h2 - mh + 1, w2 - mw + 1
        in2 = torch.from_numpy(rng.standard_normal((n, c, h2, w2)).astype(np.float32)).cuda()
        in1 = torch.from_numpy(rng.standard_normal((n, c, h1, w1)).astype(np.float32)).cuda()
        shapes = {"index": ((n, h1, w1), torch.int64), "min_ssd": ((n, h1, w1), torch.float32),
                  "pmax": ((n, h1, w1), torch.float32), "index_thr": ((n, h1, w1), torch.int64),
                  "score_thr": ((n, h1, w1), torch.float32), "soft_yx": ((n, 2, h1, w1), torch.float32),
                  "conf_marginal": ((n, h1, w1), torch.float32), "flow_full": ((n, 2, h2, w2), torch.float32)}
        bufs, out = {}, {}
        for k, (shp, dt) in shapes.items():
            numel = int(np.prod(shp))
            sentinel = -12345 if dt == torch.int64 else -12345.5
            b = torch.full((numel + 2 * G,), sentinel, dtype=dt, device="cuda")
            bufs[k] = (b, sentinel, numel)
            out[k] = b[G:G + numel].view(shp)
        dm.match_extract(in1, in2, mh, mw, canvas=(h2, w2), out=out,
                         want=("index", "min_ssd", "pmax", "index_thr", "score_thr", "soft_yx", "conf_marginal"))
        torch.cuda.synchronize()
        for k, (b, sentinel, numel) in bufs.items():
            assert bool((b[:G] == sentinel).all()) and bool((b[G + numel:] == sentinel).all()), (k, n, c, h2, w2, mh, mw)
            assert not bool((out[k] == sentinel).any()), k
        # the volume, through the C ABI with a raw pointer into a guarded buffer
        K = mh * mw
        numel = n * h1 * w1 * K
        vb = torch.full((numel + 2 * G,), -12345.5, device="cuda")
        ctx = dm.default_context()
        p = _lib.dm_pair()
        p.in1, p.in2 = in1.data_ptr(), in2.data_ptr()
        p.n_pairs, p.channels, p.h1, p.w1, p.h2, p.w2 = n, c, h1, w1, h2, w2
        for mode in (0, 1):
            vb.fill_(-12345.5)
            ctx.use_stream(torch.cuda.current_stream().cuda_stream)
            dm.api.check(ctx._lib.dm_match_volume(ctx.handle, C.byref(p), mh, mw, mode, C.c_void_p(vb.data_ptr() + 4 * G)))
            torch.cuda.synchronize()
            assert bool((vb[:G] == -12345.5).all()) and bool((vb[G + numel:] == -12345.5).all()), ("volume", mode, n, c, h2, w2)
            assert not bool((vb[G:G + numel] == -12345.5).any())


def test_windows_beyond_the_tiled_kernel_fall_back_to_the_untiled_one(dm, oracle):
    """A 129-wide window does not fit a TMA box and a 16-channel 97x97 window does not fit the
    shared-memory ring: the call still answers (untiled kernel), bit-exact in exact mode."""
    rng = np.random.default_rng(111)
    for (c, mh, mw, h1, w1) in [(2, 5, 129, 6, 9), (16, 97, 97, 4, 7)]:
        in2 = rng.standard_normal((c, h1 + mh - 1, w1 + mw - 1)).astype(np.float32)
        in1 = (in2[:, mh // 2:mh // 2 + h1, mw // 3:mw // 3 + w1] + 0.1 * rng.standard_normal((c, h1, w1))).astype(np.float32)
        got = dm.match_extract(in1, in2, mh, mw, exact=True, want=("index", "min_ssd", "pmax"))
        vol = oracle.spatial_matching(in1, in2, mh, mw).reshape(-1, mh * mw)
        prob = oracle.neg_softmax(vol)
        idx, pmax = oracle.argmax_tie(prob, mh * mw, (math.ceil(mh / 2) - 1) * mw + math.ceil(mw / 2))
        np.testing.assert_array_equal(got["index"].reshape(-1), idx)
        np.testing.assert_array_equal(got["min_ssd"].reshape(-1), vol.min(-1))
        np.testing.assert_allclose(got["pmax"].reshape(-1), pmax, rtol=1e-4)
        np.testing.assert_array_equal(dm.match_volume(in1, in2, mh, mw, exact=True).reshape(-1, mh * mw), vol)


def test_two_contexts_do_not_lower_each_others_shared_memory_grant(dm):
    """Function attributes belong to the device: a second context that launches the same kernel with
    a smaller dynamic shared-memory size must not shrink the grant the first one relies on (round 2:
    a per-context cache did exactly that -> cudaLaunchKernel 'invalid argument' on the next call)."""
    import torch
    c1, c2 = dm.Context(0), dm.Context(0)
    g = torch.Generator(device="cuda").manual_seed(3)
    big2 = torch.randn((3, 10, 120, 400), device="cuda", generator=g)
    big1 = big2[:, :, 16:16 + 88, 16:16 + 368].contiguous()
    sm2 = torch.randn((3, 10, 200, 400), device="cuda", generator=g)
    sm1 = sm2[:, :, 4:4 + 192, 4:4 + 392].contiguous()
    want = ("index", "pmax", "score_thr")
    a = dm.match_extract(big1, big2, 33, 33, want=want, ctx=c1)          # large ring
    dm.match_extract(sm1, sm2, 9, 9, want=want, ctx=c2)                  # same kernel, small ring
    b = dm.match_extract(big1, big2, 33, 33, want=want, ctx=c1)          # must still launch
    torch.cuda.synchronize()
    for k in want:
        assert torch.equal(a[k], b[k]), k
