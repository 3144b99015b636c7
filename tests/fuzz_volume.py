"""Randomised parity sweep of the volume path (nn.SpatialMatching output, with and without the
Minus + SoftMax stages) and of the feature extractor against the oracle (run on a GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import depthmatch as dm
import oracle_lib as O
from synth import make_pair

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bad = 0
for case in range(n_cases):
    C = int(rng.choice([1, 3, 4, 7, 10, 16, 19]))
    maxh, maxw = int(rng.integers(1, 18)), int(rng.integers(1, 18))
    if case % 4 == 3:   # large windows: the shapes the strip kernel is the default for (many items per step)
        maxh, maxw = [(33, 33), (32, 32), (34, 33), (21, 40), (30, 17), (17, 65), (35, 31), (33, 34)][(case // 4) % 8]
    H2, W2 = int(rng.integers(maxh + 1, maxh + 30)), int(rng.integers(maxw + 1, maxw + 140))
    strip = bool(rng.random() < 0.5)      # the strip kernel (whole pixel streams, bulk copies) wherever it fits
    if strip:
        W2 += (-(W2 - maxw + 1)) % 4      # it needs W1 % 4 == 0
    dm.default_context().set_option("volume_kernel", 2 if strip else 0)
    in1, in2, _ = make_pair(C, H2, W2, maxh, maxw, seed=int(rng.integers(1 << 30)), noise=float(rng.choice([0, 0.3])))
    exact = bool(rng.random() < 0.5)
    vol = O.spatial_matching(in1, in2, maxh, maxw)
    got = dm.match_volume(in1, in2, maxh, maxw, exact=exact)
    errs = []
    if exact:
        if not np.array_equal(got, vol):
            errs.append("ssd exact")
    elif not np.allclose(got, vol, rtol=1e-5, atol=1e-6):
        errs.append("ssd")
    prob = O.neg_softmax(vol).reshape(vol.shape)
    gp = dm.match_volume(in1, in2, maxh, maxw, softmax=True, exact=exact)
    if not np.allclose(gp, prob, rtol=1e-4, atol=1e-9):
        errs.append("softmax")
    print("volume case %2d C=%2d win=%2dx%2d in2=%3dx%3d exact=%d strip=%d: %s"
          % (case, C, maxh, maxw, H2, W2, exact, strip, "ok" if not errs else "FAIL " + ",".join(errs)), flush=True)
    bad += bool(errs)
dm.default_context().set_option("volume_kernel", 0)
for case in range(n_cases // 2):
    n_in, n_out = int(rng.integers(1, 9)), int(rng.integers(1, 13))
    kh, kw = int(rng.integers(1, 18)), int(rng.integers(1, 18))
    h, w = int(rng.integers(kh, kh + 40)), int(rng.integers(kw, kw + 90))
    pads = tuple(int(v) for v in rng.integers(0, 4, 4))
    x = rng.standard_normal((n_in, h, w)).astype(np.float32)
    use_map = n_in >= 2 and rng.random() < 0.5
    if use_map:
        nto = int(rng.integers(1, n_in + 1))
        conv = dm.nn.SpatialConvolutionMap(dm.nn.tables.random(n_in, n_out, nto, rng), kw, kh, rng)
    else:
        conv = dm.nn.SpatialConvolution(n_in, n_out, kw, kh, rng)
    tanh = bool(rng.random() < 0.5)
    flt = dm.Filter([conv] + ([dm.nn.Tanh()] if tanh else []))
    want = O.conv_layer(x, conv.weight, conv.bias, conv.connTable, pads, tanh)
    got = flt.forward(x, pads)
    ok = got.shape == want.shape and np.allclose(got, want, rtol=1e-4, atol=1e-4 * max(1.0, float(np.abs(want).max())))
    print("filter case %2d %d->%d %dx%d in=%dx%d pads=%s map=%d tanh=%d: %s"
          % (case, n_in, n_out, kh, kw, h, w, pads, use_map, tanh, "ok" if ok else "FAIL"), flush=True)
    bad += not ok
print("failures:", bad)
sys.exit(1 if bad else 0)
