"""BASELINE.json configurations at FULL size against the CPU oracle, in the form the library picks
by itself (the dot form for these sizes) -- the comparison bench.py's headline rests on.

The oracle takes 0.6-3 s per 640x360 / 33x33 pair on the GPU box's host cores (1.7 GB of volume +
probabilities), so each case here is one pair (or one row band).  Bars: parity.py.
"""
import math

import numpy as np
import pytest

from parity import assert_parity, oracle_pair, parity_report
from synth import make_pair

pytestmark = pytest.mark.gpu

WANT = ("index", "min_ssd", "pmax", "index_thr", "score_thr")


def _pair(C, H, W, maxh, maxw, data, seed):
    """planted: frame 1 = frame 2 displaced by a smooth integer flow + N(0, sigma^2); unrelated: two
    independent N(0,1) maps (no match anywhere: flat soft-max, many near-ties, the hard case)."""
    if data == "unrelated":
        rng = np.random.default_rng(seed)
        in2 = rng.standard_normal((C, H, W), dtype=np.float32)
        in1 = rng.standard_normal((C, H - maxh + 1, W - maxw + 1), dtype=np.float32)
        return in1, in2
    in1, in2, _ = make_pair(C, H, W, maxh, maxw, seed=seed, noise=float(data))
    return in1, in2


def _norm_sum(in1, in2):
    return float((in1.astype(np.float64) ** 2).sum(0).max() + (in2.astype(np.float64) ** 2).sum(0).max())


def _dot_limit(C):
    return 1e-4 / ((C + 2) * 2.0 ** -24)


def _check(dm, oracle, in1, in2, maxh, maxw, canvas, what, expect_dot=True):
    ctx = dm.default_context()
    l0 = ctx.launch_count()
    got = dm.match_extract(in1, in2, maxh, maxw, canvas=canvas, want=WANT)
    launches = ctx.launch_count() - l0
    rescored, exact_pass = ctx.last_counts()
    if expect_dot:
        # the library's rule: large call and max|a|^2 + max|b|^2 under the bound -> dot form
        # (2 norm launches + twin sweep + rescore + exact pass = 6 launches)
        assert _norm_sum(in1, in2) < _dot_limit(in1.shape[0]) and launches == 6, (launches, _norm_sum(in1, in2))
    want = oracle_pair(oracle, in1, in2, maxh, maxw, canvas=canvas)
    rep = parity_report(got, want)
    rep.update(rescored=rescored, exact_pass=exact_pass, launches=launches)
    print("\n[parity] %s: %r" % (what, rep))
    assert_parity(rep, what)
    return rep


@pytest.mark.parametrize("data", ["0.05", "0.3", "unrelated"])
def test_north_full_pair_vs_oracle(dm, oracle, data):
    """BASELINE north: 640x360 feature maps, C = 10, 33x33 window; the benchmark's own pair (seed
    1234, sigma 0.05), a noisy one and two unrelated frames."""
    in1, in2 = _pair(10, 360, 640, 33, 33, data, 1234)
    rep = _check(dm, oracle, in1, in2, 33, 33, (360, 640), "north/" + data)
    assert rep["pixels"] == 328 * 608


def test_c2_pairs_of_the_batch_vs_oracle(dm, oracle):
    """configs[1]: 320x180, 33x33, batch 64.  The whole batch runs in one call (dot form); pairs 0,
    31 and 63 are compared with the oracle."""
    import torch
    N = 64
    pairs = [_pair(10, 180, 320, 33, 33, "0.05" if n % 3 else "unrelated", 300 + n) for n in range(N)]
    in1 = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
    in2 = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    got = dm.match_extract(in1, in2, 33, 33, canvas=(180, 320), want=WANT)
    torch.cuda.synchronize()
    for n in (0, 31, 63):
        want = oracle_pair(oracle, pairs[n][0], pairs[n][1], 33, 33, canvas=(180, 320))
        rep = parity_report({k: v[n].cpu().numpy() for k, v in got.items()}, want)
        print("\n[parity] c2 pair %d: %r" % (n, rep))
        assert_parity(rep, "c2 pair %d" % n)


def test_c5_row_band_vs_oracle(dm, oracle):
    """configs[4]: 1920x1080, 65x65 window.  32 output rows of row band 3 of 8 (what rank 3 computes,
    halo included) against the oracle on the same rows: 1 GB of volume on the host."""
    from depthmatch import parallel
    maxh = maxw = 65
    C, H, W = 10, 1080, 1920
    H1 = H - maxh + 1
    y0, y1, _ = parallel.row_bands(H1, 8, maxh)[3]
    rng = np.random.default_rng(55)
    rows = 32
    in2 = rng.standard_normal((C, rows + maxh - 1, W), dtype=np.float32)   # band rows y0+40 .. of frame 2
    W1 = W - maxw + 1
    fy, fx = 7, -11
    in1 = in2[:, 32 + fy:32 + fy + rows, 32 + fx:32 + fx + W1] + 0.3 * rng.standard_normal((C, rows, W1), dtype=np.float32)
    in1 = np.ascontiguousarray(in1)
    in1[:, :, W1 // 2:] = rng.standard_normal((C, rows, W1 - W1 // 2), dtype=np.float32)   # right half: no match
    rep = _check(dm, oracle, in1, in2, maxh, maxw, None, "c5 band rows")
    assert rep["pixels"] == rows * W1 and y1 - y0 == 127


@pytest.mark.parametrize("scale,expect_dot", [(1.28, True), (1.6, False)])
def test_dot_form_at_its_norm_limit(dm, oracle, scale, expect_dot):
    """ADVICE r1 / VERDICT r1 weak-2: the dot form's error grows with |a|^2 + |b|^2.  Features scaled
    so that the largest norm sum sits just under the bound the library derives from the 1e-4 score
    bar (139.8 for 10 channels) must still meet every bar on hard data (unrelated frames in half of
    the image); just above it the device-side switch runs the difference form."""
    maxh = maxw = 33
    in1, in2, _ = make_pair(10, 200, 400, maxh, maxw, seed=77, noise=0.3)
    rng = np.random.default_rng(78)
    in1[:, :, in1.shape[2] // 2:] = rng.standard_normal(in1[:, :, in1.shape[2] // 2:].shape, dtype=np.float32)
    # bring the largest norm sum to `scale`/1.28 * 0.97 of the limit
    target = 0.97 * _dot_limit(10) * (scale / 1.28) ** 2
    f = math.sqrt(target / _norm_sum(in1, in2))
    in1, in2 = (in1 * f).astype(np.float32), (in2 * f).astype(np.float32)
    ns = _norm_sum(in1, in2)
    assert (ns < _dot_limit(10)) == expect_dot
    rep = _check(dm, oracle, in1, in2, maxh, maxw, None, "dot limit x%.2f (norm sum %.1f)" % (scale, ns), expect_dot=False)
    if expect_dot:
        dif = dm.match_extract(in1, in2, maxh, maxw, want=("pmax",), diff_form=True)
        dot = dm.match_extract(in1, in2, maxh, maxw, want=("pmax",))
        assert not np.array_equal(dif["pmax"], dot["pmax"])       # the dot kernel really ran
    else:
        dif = dm.match_extract(in1, in2, maxh, maxw, want=("index", "pmax"), diff_form=True)
        dot = dm.match_extract(in1, in2, maxh, maxw, want=("index", "pmax"))
        for k in dif:
            np.testing.assert_array_equal(dif[k], dot[k])         # fell back, bit for bit


def test_hard_data_33x33_small_sizes_forced_dot(dm, oracle):
    """Hard data where the oracle is cheap, dot form forced: sigma 0.3 / 1.0 and unrelated frames,
    two tiles wide, with a flat (all-tie) block."""
    ctx = dm.default_context()
    ctx.set_option("ssd_form", "dot")
    for data, seed in (("0.3", 1), ("1.0", 2), ("unrelated", 3)):
        in1, in2 = _pair(10, 70, 190, 33, 33, data, seed)
        in1[:, 5:12, 20:60] = 0.5
        in2[:, 5:44, 20:92] = 0.5
        got = dm.match_extract(in1, in2, 33, 33, canvas=(70, 190), want=WANT + ("soft_yx",))
        want = oracle_pair(oracle, in1, in2, 33, 33, canvas=(70, 190))
        rep = parity_report(got, want)
        rep["rescored"] = ctx.last_counts()[0]
        print("\n[parity] 33x33 %s forced dot: %r" % (data, rep))
        assert_parity(rep, data)
        assert rep["rescored"] >= 7 * 40      # the flat block cannot be ordered by the dot form
        assert (got["index"][5:12, 20:60] == want["middle"]).all()


@pytest.mark.parametrize("shape,data", [((10, 70, 190, 33, 33), "unrelated"), ((10, 100, 300, 17, 24), "0.3"),
                                        ((4, 60, 140, 9, 9), "1.0"), ((10, 360, 640, 33, 33), "0.05")])
def test_two_row_sweep_variant_vs_oracle(dm, oracle, shape, data):
    """The opt-in two-rows-per-warp dot sweep (match_sweep2.cuh, option sweep = 2: measured slower than
    the default on B200, kept as the record of that experiment) meets the same bars, on both of its
    epilogues (scores and winner-take-all)."""
    C, H, W, maxh, maxw = shape
    ctx = dm.default_context()
    ctx.set_option("sweep", "2")
    ctx.set_option("ssd_form", "dot")
    try:
        in1, in2 = _pair(C, H, W, maxh, maxw, data, 91)
        in1[:, 3:9, 10:40] = 0.25
        in2[:, 3:9 + maxh - 1, 10:40 + maxw - 1] = 0.25       # a flat block: all-tie pixels go to the rescore
        want = oracle_pair(oracle, in1, in2, maxh, maxw, canvas=(H, W))
        got = dm.match_extract(in1, in2, maxh, maxw, canvas=(H, W), want=WANT)
        rep = parity_report(got, want)
        rep["rescored"] = ctx.last_counts()[0]
        print("\n[parity] two-row sweep %r %s: %r" % (shape, data, rep))
        assert_parity(rep, "two-row sweep scores")
        assert rep["rescored"] >= 6 * 30
        wta = dm.match_extract(in1, in2, maxh, maxw, canvas=(H, W), want=("index", "min_ssd"))
        assert_parity(parity_report(wta, want), "two-row sweep wta")
        np.testing.assert_array_equal(wta["index"], got["index"])
    finally:
        ctx.set_option("sweep", "0")
        ctx.set_option("ssd_form", "auto")


@pytest.mark.parametrize("kind,layers,shape,pads", [
    ("filter", [[3, 5, 5, 8], [4, 16, 16, 10]], (2, 3, 180, 320), (0, 0, 0, 0)),   # c1: full layer, tanh, connection table
    ("filter", [[3, 5, 5, 10]], (1, 3, 90, 150), (2, 2, 2, 2)),                     # c3's prefilter with zero padding
    ("radial", [[3, 1, 17, 5], [5, 17, 1, 10]], (2, 3, 100, 216), (0, 0, 0, 0)),    # c4's radial net: 1 x 17 then 17 x 1
    ("radial", [[1, 7, 9, 4]], (1, 1, 40, 300), (0, 0, 0, 0)),                      # one input plane: the second K half is empty
])
def test_tensor_core_filter_layers_vs_default_kernel(dm, kind, layers, shape, pads):
    """The opt-in tcgen05 convolution (filter_tc.cu, option conv = 2: kind::tf32 with a three-term hi / lo
    split, accumulators in TMEM) against the default CUDA-core kernel -- itself held to 1e-4 of the
    fp32 oracle by test_gpu_parity.py -- at the feature bar."""
    rng = np.random.default_rng(5)
    ctx = dm.default_context()
    flt = dm.getFilter(dm.Geometry(layers=layers), rng) if kind == "filter" else dm.getRadialFilter(dict(layers=layers), rng)
    x = rng.standard_normal(shape).astype(np.float32)
    base = flt.forward(x, pads)
    ctx.set_option("conv", "2")
    try:
        l0 = ctx.launch_count()
        got = flt.forward(x, pads)
        assert ctx.launch_count() - l0 == len(layers)
    finally:
        ctx.set_option("conv", "0")
    scale = max(float(np.abs(base).max()), 1.0)
    assert float(np.abs(got - base).max()) <= 1e-4 * scale, float(np.abs(got - base).max())
    assert not np.array_equal(got, base)          # a different kernel really ran


@pytest.mark.parametrize("data", ["0.3", "unrelated"])
def test_softmax_volume_of_a_large_call_vs_oracle(dm, oracle, data):
    """dm_match_volume(NEG_SOFTMAX) on a call large enough for the dot form (both of its sweeps: the
    statistics and the volume): every probability within 1e-4 of Minus + SoftMax on the CPU."""
    maxh = maxw = 33
    in1, in2 = _pair(10, 200, 400, maxh, maxw, data, 61)
    ctx = dm.default_context()
    l0 = ctx.launch_count()
    got = dm.match_volume(in1, in2, maxh, maxw, softmax=True)
    assert ctx.launch_count() - l0 == 4            # 2 norm passes + twin strip kernels (soft-max in the staging buffer)
    ctx.set_option("volume_kernel", 1)
    l0 = ctx.launch_count()
    two = dm.match_volume(in1, in2, maxh, maxw, softmax=True)
    ctx.set_option("volume_kernel", 0)
    assert ctx.launch_count() - l0 == 6            # tiled kernel: 2 norm passes + twin statistics + twin volume sweeps
    np.testing.assert_allclose(got, two, rtol=2e-5, atol=1e-12)
    want = oracle.neg_softmax(oracle.spatial_matching(in1, in2, maxh, maxw))
    np.testing.assert_allclose(got, want.reshape(got.shape), rtol=1e-4, atol=1e-12)
    dif = dm.match_volume(in1 * 4, in2 * 4, maxh, maxw, softmax=True)      # norms beyond the bound: difference form
    want = oracle.neg_softmax(oracle.spatial_matching(in1 * 4, in2 * 4, maxh, maxw))
    np.testing.assert_allclose(dif, want.reshape(dif.shape), rtol=1e-4, atol=1e-12)


def test_volume_mode_north_full_pair_vs_oracle(dm, oracle):
    """Volume mode (nn.SpatialMatching's output, then Minus + SoftMax) at the benchmarked size: one whole
    640x360 / 33x33 pair, 869 MB per volume, through the strip kernel (bulk-copied pixel streams; soft-max taken in
    its staging buffer) -- every entry against the oracle: SSD bit-exact in exact mode, 1e-4 otherwise."""
    import torch
    maxh = maxw = 33
    in1, in2 = _pair(10, 360, 640, maxh, maxw, "0.05", 1234)
    vol = oracle.spatial_matching(in1, in2, maxh, maxw)
    t1, t2 = torch.from_numpy(in1).cuda(), torch.from_numpy(in2).cuda()
    ctx = dm.default_context()
    l0 = ctx.launch_count()
    got = dm.match_volume(t1, t2, maxh, maxw, exact=True)
    assert ctx.launch_count() - l0 == 1                      # one strip kernel, nothing else
    assert torch.equal(got.cpu(), torch.from_numpy(vol).view(got.shape))
    got = dm.match_volume(t1, t2, maxh, maxw)
    np.testing.assert_allclose(got.cpu().numpy().reshape(vol.shape), vol, rtol=1e-4, atol=1e-5)
    del got
    prob = oracle.neg_softmax(vol)
    del vol
    l0 = ctx.launch_count()
    gp = dm.match_volume(t1, t2, maxh, maxw, softmax=True)
    assert ctx.launch_count() - l0 == 4                      # norm pre-pass (2) + twin strip kernels: no statistics sweep
    np.testing.assert_allclose(gp.cpu().numpy().reshape(prob.shape), prob, rtol=1e-4, atol=1e-12)
