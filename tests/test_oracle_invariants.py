"""The reference's own behavioural tests restated against the
oracle.  These are what pin the parts of the oracle whose arithmetic lives in un-vendored
Torch7 code (parity otherwise unpinned)."""
import math

import numpy as np
import pytest

from synth import make_pair


# T1: tests/test_multiscale.lua:149-166 -- SpatialMatching argmin == brute-force SSD argmin
def test_matching_argmin_is_bruteforce_ssd(oracle):
    rng = np.random.default_rng(0)
    C, maxh, maxw, H1, W1 = 5, 6, 7, 9, 10
    in2 = rng.standard_normal((C, H1 + maxh - 1, W1 + maxw - 1)).astype(np.float32)
    in1 = rng.standard_normal((C, H1, W1)).astype(np.float32)
    vol = oracle.spatial_matching(in1, in2, maxh, maxw)
    for y in range(H1):
        for x in range(W1):
            best, bi, bj = 1e25, 0, 0
            for i in range(maxh):
                for j in range(maxw):
                    s = float(((in1[:, y, x] - in2[:, y + i, x + j]) ** 2).sum())
                    if s < best:
                        best, bi, bj = s, i, j
            m = int(np.argmin(vol[y, x].reshape(-1)))
            assert (m // maxw, m % maxw) == (bi, bj)
            np.testing.assert_allclose(vol[y, x, bi, bj], best, rtol=1e-5)


# T2: cartesian_groundtruth_cc_testme (radial/radial_opticalflow_groundtruth.lua:170-210):
# integer warps are recovered exactly by matching + argmin + zero-flow tie rule
@pytest.mark.parametrize("maxh,maxw", [(12, 15), (17, 15), (17, 17)])
def test_integer_warp_known_answer(oracle, maxh, maxw):
    in1, in2, flow = make_pair(30, 32 + maxh - 1, 42 + maxw - 1, maxh, maxw, seed=5, noise=0.0)
    vol = oracle.spatial_matching(in1, in2, maxh, maxw)
    K = maxh * maxw
    middle = math.ceil(maxw / 2) + maxw * (math.ceil(maxh / 2) - 1)
    idx, _ = oracle.argmin_tie(vol, K, middle)
    idx = idx.reshape(in1.shape[1:])
    fy = (idx - 1) // maxw - (math.ceil(maxh / 2) - 1)
    fx = (idx - 1) % maxw - (math.ceil(maxw / 2) - 1)
    np.testing.assert_array_equal(fy, flow[0])
    np.testing.assert_array_equal(fx, flow[1])
    # and through the softmax + argmax + canvas path of processOutput
    prob = oracle.neg_softmax(vol)
    idx2, _ = oracle.argmax_tie(prob, K, middle)
    full = oracle.flow_canvas(idx2, in1.shape[1], in1.shape[2], maxh, maxw, in2.shape[1], in2.shape[2])
    hoff, woff = (in2.shape[1] - in1.shape[1]) // 2, (in2.shape[2] - in1.shape[2]) // 2
    np.testing.assert_array_equal(full[0, hoff:hoff + in1.shape[1], woff:woff + in1.shape[2]], flow[0])
    np.testing.assert_array_equal(full[1, hoff:hoff + in1.shape[1], woff:woff + in1.shape[2]], flow[1])
    assert full[:, :hoff].sum() == 0 and full[:, :, :woff].sum() == 0


# T3: tests/test_multiscale.lua:57-80 -- ring index round trips; L = 112 / 160
@pytest.mark.parametrize("ratios,L", [([1, 2], 112), ([1, 2, 4], 160)])
def test_ring_index_round_trip(oracle, ratios, L):
    maxh = maxw = 8
    assert oracle.multiscale_length(maxh, maxw, ratios) == L
    for i in range(1, L + 1):
        rc, y, x = oracle.x2yx_multi_number(maxh, maxw, ratios, i)
        assert rc == 0
        assert oracle.yx2x_multi(maxh, maxw, ratios, y, x) == i
    mh, mw = maxh * ratios[-1], maxw * ratios[-1]
    for i in range(-math.ceil(mh / 2) + 1, mh // 2 + 1):
        for j in range(-math.ceil(mw / 2) + 1, mw // 2 + 1):
            rc, y, x = oracle.x2yx_multi_number(maxh, maxw, ratios, oracle.yx2x_multi(maxh, maxw, ratios, i, j))
            tol = 1
            for r in ratios:
                if abs(i) < maxh * r and abs(j) < maxw * r:
                    tol = r
            assert rc == 0 and abs(y - i) < tol and abs(x - j) < tol
    assert oracle.yx2x_multi(maxh, maxw, ratios, 0, 0) == 28  # getMiddleIndex


# T4: tests/test_multiscale.lua:169-193 without the stale /n -- cascade = nearest-upsampled crop sum
def test_cascade_is_nearest_upsampled_sum(oracle):
    rng = np.random.default_rng(1)
    ratios = [1, 2, 4]
    inp = rng.random((3, 6, 8, 8)).astype(np.float32)
    out = oracle.cascade_add(inp, ratios)
    np.testing.assert_array_equal(out[2], inp[2])
    cy = cx = 4
    for i in range(3):
        s = np.zeros((6, 8, 8), np.float64)
        for ii in range(-cy + 1, cy + 1):
            for jj in range(-cx + 1, cx + 1):
                for j in range(i, 3):
                    r = ratios[j] / ratios[i]
                    s[:, ii + cy - 1, jj + cx - 1] += inp[j][:, math.ceil(ii / r) + cy - 1, math.ceil(jj / r) + cx - 1]
        np.testing.assert_allclose(out[i], s, rtol=1e-6)


# T5: tests/test_multiscale.lua:195-214 -- ring layout [top | left | right | bottom]
def test_ring_layout(oracle):
    rng = np.random.default_rng(2)
    ratios = [1, 2]
    casc = rng.random((2, 3, 8, 8)).astype(np.float32)
    vec = oracle.ring_join(casc, ratios)
    assert vec.shape == (3, 112)
    np.testing.assert_array_equal(vec[:, :64], casc[0].reshape(3, 64))
    d, li = 2, 4
    ring = vec[:, 64:]
    block = np.zeros((3, 8, 8), np.float32)
    block[:, :d] = ring[:, : d * 8].reshape(3, d, 8)
    block[:, d:d + li, :d] = ring[:, d * 8: d * 8 + li * d].reshape(3, li, d)
    block[:, d:d + li, d + li:] = ring[:, d * 8 + li * d: d * 8 + 2 * li * d].reshape(3, li, d)
    block[:, d + li:] = ring[:, d * 8 + 2 * li * d:].reshape(3, d, 8)
    ref = casc[1].copy()
    ref[:, d:d + li, d:d + li] = 0
    np.testing.assert_array_equal(block, ref)


# T6: cartesian2polar_testme (radial/cartesian2polar.lua:95-106) -- cart -> polar -> cart
def test_polar_round_trip(oracle):
    h, w = 116, 226
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    img = (np.sin(xx / 17.0) + np.cos(yy / 11.0) + 0.01 * xx).astype(np.float32)[None]
    rmax = min(h // 2, w // 2) - 1
    m = oracle.c2p_mask(400, 250, w / 2, h / 2, 0, 0, rmax)
    pol = oracle.warp_bilinear(img, m)
    m2 = oracle.p2c_mask(400, 250, w, h, w / 2, h / 2, rmax)
    back = oracle.warp_bilinear(pol, m2)
    r = np.hypot(xx - w / 2, yy - h / 2)
    inside = (r < rmax - 2) & (r > 3)
    assert np.abs(back[0] - img[0])[inside].max() < 0.05


def test_c2p_mask_padding_is_circular(oracle):
    m = oracle.c2p_mask(40, 30, 100.3, 80.7, 3, 4, 55.0, 1.0)
    core = m[:, :, 3:43]
    np.testing.assert_array_equal(m[:, :, :3], core[:, :, -3:])
    np.testing.assert_array_equal(m[:, :, 43:], core[:, :, :4])
    assert core[0, 0, 0] == np.float32(80.7) and core[1, 0, 0] == np.float32(100.3)


def test_softmax_rows_sum_to_one_and_th_approx_is_close(oracle):
    rng = np.random.default_rng(4)
    vol = (rng.random((50, 64)) * 12).astype(np.float32)
    p = oracle.neg_softmax(vol)
    np.testing.assert_allclose(p.sum(-1), 1.0, rtol=1e-6)
    pa = oracle.neg_softmax(vol, exp_mode=1)
    # the 2012 polynomial exp is a ~1e-3 approximation: same argmax, close scores
    assert np.array_equal(p.argmax(-1), pa.argmax(-1))
    assert np.abs(p - pa).max() < 5e-3


def test_soft_mean_and_marginal(oracle):
    rng = np.random.default_rng(5)
    p = rng.random((20, 5 * 7)).astype(np.float32)
    p /= p.sum(-1, keepdims=True)
    ym, xm = oracle.soft_mean(p, 5, 7)
    rows, cols = np.meshgrid(np.arange(1, 6), np.arange(1, 8), indexing="ij")
    np.testing.assert_allclose(ym, (p * rows.reshape(-1)).sum(-1), rtol=1e-5)
    np.testing.assert_allclose(xm, (p * cols.reshape(-1)).sum(-1), rtol=1e-5)
    np.testing.assert_allclose(oracle.marginal_x(p, 5, 7), p.reshape(20, 5, 7).sum(-1), rtol=1e-5)


def test_flow2depth_rules(oracle):
    flow = np.full((40, 50), 2.0, np.float32)
    flow[5, 5] = 0.05
    depth, conf = oracle.flow2depth(flow, 25.0, 20.0, 99.0)
    assert conf[20, 25] == 0 and depth[20, 25] == 0          # within 10 px of the epipole
    assert depth[5, 5] == 99.0                               # flow < 0.1 -> infinity
    np.testing.assert_allclose(depth[0, 0], math.hypot(25, 20) / 2.0, rtol=1e-6)


# ---- T7/T8: the feature extractor (getFilter) ------------------------------------------------
def _dense_from_map(conn, wm, n_in, n_out):
    wd = np.zeros((n_out, n_in) + wm.shape[1:], np.float32)
    for e, (f, t) in enumerate(conn):
        wd[t - 1, f - 1] += wm[e]
    return wd


def test_conv_layer_equals_an_independent_conv2d(oracle):
    """T7: nn.SpatialConvolution / SpatialConvolutionMap + SpatialZeroPadding + Tanh against
    torch's conv2d (an independent implementation of the same valid cross-correlation)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(11)
    x = rng.standard_normal((3, 23, 31)).astype(np.float32)
    w = rng.standard_normal((8, 3, 5, 5)).astype(np.float32) * 0.2
    b = rng.standard_normal(8).astype(np.float32)
    got = oracle.conv_layer(x, w, b, pads=(1, 2, 3, 0), tanh=True)
    want = torch.tanh(F.conv2d(F.pad(torch.from_numpy(x)[None], (1, 2, 3, 0)), torch.from_numpy(w),
                               torch.from_numpy(b)))[0].numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    conn = np.array([[f + 1, t + 1] for t in range(10) for f in rng.permutation(8)[:4]], np.int32)
    wm = rng.standard_normal((40, 16, 16)).astype(np.float32) * 0.05
    bm = rng.standard_normal(10).astype(np.float32)
    x8 = rng.standard_normal((8, 30, 41)).astype(np.float32)
    got = oracle.conv_layer(x8, wm, bm, conn)
    want = F.conv2d(torch.from_numpy(x8)[None], torch.from_numpy(_dense_from_map(conn, wm, 8, 10)),
                    torch.from_numpy(bm))[0].numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)


def test_patch_unfolding_filter_makes_matching_a_patch_ssd(oracle):
    """T8 (tests/test_multiscale.lua:44-55,135-166): with identity 'patch-unfolding' weights
    (output plane (c,i,j) copies input pixel (c, y+i, x+j)), SpatialMatching on the features is
    the brute-force SSD between k x k patches."""
    rng = np.random.default_rng(12)
    k, c = 3, 2
    w = np.zeros((c * k * k, c, k, k), np.float32)
    for ci in range(c):
        for i in range(k):
            for j in range(k):
                w[(ci * k + i) * k + j, ci, i, j] = 1
    b = np.zeros(c * k * k, np.float32)
    img2 = rng.standard_normal((c, 16, 18)).astype(np.float32)
    img1 = (np.roll(img2, (-1, -2), (1, 2)) + 0.01 * rng.standard_normal(img2.shape)).astype(np.float32)
    f1 = oracle.conv_layer(img1[:, 1:-1, 1:-1], w, b)      # 12 x 14
    f2 = oracle.conv_layer(img2, w, b)                      # 14 x 16
    vol = oracle.spatial_matching(f1, f2, 3, 3)
    h1, w1 = f1.shape[1:]
    for (y, x) in ((0, 0), (5, 7), (h1 - 1, w1 - 1)):
        p1 = img1[:, 1 + y:1 + y + k, 1 + x:1 + x + k]
        for dy in range(3):
            for dx in range(3):
                p2 = img2[:, y + dy:y + dy + k, x + dx:x + dx + k]
                np.testing.assert_allclose(vol[y, x, dy, dx], ((p1 - p2) ** 2).sum(), rtol=1e-5, atol=1e-6)
