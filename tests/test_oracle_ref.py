"""Pins the oracle restatement against the reference's OWN native code compiled as-is
(oracle/_ref/libdm_ref.so: version2/extract_output.cpp and x2yxMulti2.c from
/root/reference, see oracle/Makefile).  On the GPU box /root/reference does not exist but the
built oracle/_ref travels with the snapshot; if it is absent these tests skip and the
committed golden vectors (tests/golden/, generated from the same _ref) still pin the oracle.
"""
import numpy as np
import pytest


def _need_ref(oracle):
    if oracle.ref() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")


@pytest.mark.parametrize("threshold", [0.11, 0.21, 0.0, -0.5, 0.199999, 0.2, 0.9])
@pytest.mark.parametrize("n", [1, 5, 9, 64, 289, 1089])
def test_extract_output_matches_reference_source(oracle, threshold, n):
    _need_ref(oracle)
    rng = np.random.default_rng(n * 1000 + int(threshold * 100) + 7)
    inp = rng.random((6, 11, n), dtype=np.float32) * (0.5 if n < 64 else 0.24)
    inp[0, 0] = 0.0                      # nothing above a positive threshold
    inp[1, 1] = inp[1, 1, 0]             # all equal: ties go through the sorting network
    inp[2, 2, : min(n, 9)] = 0.3         # more than M qualifiers, scan-order cut-off
    r0 = rng.integers(-5, 5, (6, 11))
    s0 = rng.random((6, 11)).astype(np.float32)
    a = oracle.extract_output(inp, threshold, r0, s0, "oracle")
    b = oracle.extract_output(inp, threshold, r0, s0, "ref")
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


@pytest.mark.parametrize("threshold,acc", [(0.11, 0.5), (0.21, 1.0), (0.05, 0.0)])
def test_extract_output_marginalized_matches_reference_source(oracle, threshold, acc):
    _need_ref(oracle)
    rng = np.random.default_rng(3)
    inp = rng.random((9, 7, 33), dtype=np.float32) * 0.3
    r0 = rng.integers(0, 9, (9, 7))
    a = oracle.extract_output_marginalized(inp, threshold, acc, r0, "oracle")
    b = oracle.extract_output_marginalized(inp, threshold, acc, r0, "ref")
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


@pytest.mark.parametrize("maxh,maxw,ratios", [(8, 8, [1, 2]), (8, 8, [1, 2, 4]), (16, 16, [1, 2, 4]),
                                             (9, 7, [1, 3]), (8, 8, [1]), (12, 8, [1, 2])])
def test_x2yxmulti2_bugcompat_matches_reference_source(oracle, maxh, maxw, ratios):
    _need_ref(oracle)
    L = oracle.multiscale_length(maxh, maxw, ratios)
    x = np.arange(-2, L + 40 + (L % 2), dtype=np.int64).reshape(2, -1)
    a = oracle.x2yx_multi2_bugcompat(x, maxh, maxw, ratios, "oracle")
    b = oracle.x2yx_multi2_bugcompat(x, maxh, maxw, ratios, "ref")
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_reference_c_decode_diverges_from_lua_spec(oracle):
    """Documents SURVEY 8a-11: the shipped C is not the Lua spec (e.g. index 64 of an 8x8 window)."""
    ry, rx = oracle.x2yx_multi2_bugcompat(np.array([[64]]), 8, 8, [1, 2])
    rc, sy, sx = oracle.x2yx_multi_number(8, 8, [1, 2], 64)
    assert rc == 0 and (sy, sx) == (4, 4)
    assert (int(ry[0, 0]), int(rx[0, 0])) != (sy, sx)


# ---- the reference's `inline.load` C bodies, extracted verbatim (oracle/extract_inline.py) ----
@pytest.mark.parametrize("method", ["med", "max"])
@pytest.mark.parametrize("k", [3, 5])
def test_post_process_image_matches_reference_inline_c(oracle, method, k):
    _need_ref(oracle)
    rng = np.random.default_rng(k)
    # the mode filter indexes a 16 x 16 histogram without a range check: flows span < 16 values
    flow = np.clip(np.rint(rng.normal(0, 3, (2, 31, 44))), -7, 8).astype(np.float32)
    if method == "med":
        flow = flow + rng.random((2, 31, 44)).astype(np.float32)
    mask = (rng.random((31, 44)) > 0.3).astype(np.float32)
    mask[10:14, 20:26] = 0  # windows without a single masked pixel
    a = oracle.post_process_image(flow, mask, k, method, "oracle")
    b = oracle.post_process_image(flow, mask, k, method, "ref")
    np.testing.assert_array_equal(a, b)


def test_enlarge_mask_matches_reference_inline_c(oracle):
    _need_ref(oracle)
    rng = np.random.default_rng(1)
    for ix, iy in ((3, 2), (1, 1), (7, 9), (0, 4)):
        mask = (rng.random((25, 33)) > 0.4).astype(np.float32)
        mask[5] = 0
        mask[:, 7] = 0
        np.testing.assert_array_equal(oracle.enlarge_mask(mask, ix, iy, "oracle"),
                                      oracle.enlarge_mask(mask, ix, iy, "ref"))


def test_radial_depth_matches_reference_inline_c(oracle):
    _need_ref(oracle)
    rng = np.random.default_rng(2)
    flow = (rng.normal(0, 2, (2, 40, 60)) * (rng.random((2, 40, 60)) > 0.2)).astype(np.float32)
    a = oracle.radial_depth(flow, 19.5, 31.25, 30.0, "oracle")
    b = oracle.radial_depth(flow, 19.5, 31.25, 30.0, "ref")
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_polar_luts_and_flow2depth_match_reference_inline_c(oracle):
    _need_ref(oracle)
    for (wd, hd, xc, yc, rmax, alpha) in ((48, 50, 40.09, 21.56, 45.0, 1.0), (400, 400, 320.73, 172.48, 367.0, 1.0),
                                          (64, 32, 30.5, 17.25, 33.0, 1.5)):
        np.testing.assert_array_equal(oracle.c2p_mask(wd, hd, xc, yc, 0, 0, rmax, alpha),
                                      oracle.ref_c2p_mask(wd, hd, xc, yc, rmax, alpha))
        np.testing.assert_array_equal(oracle.p2c_mask(wd, hd, 80, 45, xc, yc, rmax, alpha),
                                      oracle.ref_p2c_mask(wd, hd, 80, 45, xc, yc, rmax, alpha))
    rng = np.random.default_rng(3)
    flow = (rng.random((37, 41)) * 2).astype(np.float32)
    a = oracle.flow2depth(flow, 20.3, 18.9, 55.5)
    b = oracle.ref_flow2depth(flow, 20.3, 18.9, 55.5)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


@pytest.mark.parametrize("seed,shape", [(1, (24, 40)), (2, (7, 9)), (3, (36, 63))])
def test_drone_depth_from_xflow_matches_reference_source(oracle, seed, shape):
    """orc_depth_from_xflow against ARdroneAPI::computeDepthMapFromFlow compiled from
    /root/reference/ardrone/ardrone_api.cpp:99-140 (behind a cv::Mat_<float> stand-in).  Flows stay in
    the range the reference's 20-bin histogram can hold (rounded flow in [-8, 11])."""
    _need_ref(oracle)
    rng = np.random.default_rng(seed)
    h, w = shape
    xflow = rng.uniform(-7.4, 10.4, (h, w)).astype(np.float32)
    xflow[rng.random((h, w)) < 0.2] = 0.0
    mask = (rng.random((h, w)) < 0.7).astype(np.float32)
    mask[rng.random((h, w)) < 0.1] = 0.4          # non-zero but not confident
    da, ca = oracle.depth_from_xflow(xflow, mask, 0.37, "oracle")
    db, cb = oracle.depth_from_xflow(xflow, mask, 0.37, "ref")
    np.testing.assert_array_equal(ca, cb)
    np.testing.assert_array_equal(da[cb > 0], db[cb > 0])
    assert (cb > 0).sum() > 10
