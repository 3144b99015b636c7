"""Randomised parity sweep of the remaining entry points against the oracle (run on a GPU box):
multiscale model, stand-alone extractOutput / cascade / x2yxMulti, post-processing, warps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import depthmatch as dm
import oracle_lib as O

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bad = 0


def report(name, desc, ok):
    global bad
    print("%-12s %-60s %s" % (name, desc, "ok" if ok else "FAIL"), flush=True)
    bad += not ok


for case in range(n_cases):
    # ---- multiscale model (geometry constraints of getModelMultiscale: maxh*(rmax - r) even)
    ratios = [[1, 2], [1, 2, 4], [1, 3], [1, 2, 6]][int(rng.integers(4))]
    maxh = maxw = int(rng.choice([4, 6, 8, 12]))
    rmax = ratios[-1]
    H, W = rmax * int(rng.integers(2, 7)), rmax * int(rng.integers(2, 9))
    C = int(rng.choice([1, 3, 10]))
    f1s, f2s = [], []
    for r in ratios:
        h, w = H // r, W // r
        f2 = rng.standard_normal((C, h + maxh - 1, w + maxw - 1)).astype(np.float32)
        o = maxh // 2 - 1
        f1 = f2[:, o:o + h, o:o + w] + 0.5 * rng.standard_normal((C, h, w)).astype(np.float32)
        f1s.append(np.ascontiguousarray(f1))
        f2s.append(f2)
    g = dm.Geometry(maxh=maxh, maxw=maxw, ratios=ratios, multiscale=True, hImg=H, wImg=W, output_extraction_method="max")
    ok = True
    try:
        out = dm.getModelMultiscale(g, True, True).forward(list(zip(f1s, f2s)))
        K = maxh * maxw
        per = []
        for (f1, f2, r) in zip(f1s, f2s, ratios):
            prob = O.neg_softmax(O.spatial_matching(f1, f2, maxh, maxw)).reshape(f1.shape[1], f1.shape[2], K)
            per.append(O.upsample_nearest_rows(prob, r).reshape(-1, maxh, maxw))
        vec = O.ring_join(O.cascade_add(np.stack(per), ratios), ratios)
        middle = O.yx2x_multi(maxh, maxw, ratios, 0, 0)
        idx, _ = O.argmax_tie(vec, vec.shape[1], middle)
        tie = O.top2_relgap(vec, vec.shape[1]) < 1e-4
        got = np.asarray(out["index"]).reshape(-1)
        ok = not ((got != idx) & ~tie).any()
        if ok:
            for i in np.nonzero(~tie)[0][:200]:
                _, fy, fx = O.x2yx_multi_number(maxh, maxw, ratios, int(idx[i]))
                ok &= int(np.asarray(out["flow_y"]).reshape(-1)[i]) == fy and int(np.asarray(out["flow_x"]).reshape(-1)[i]) == fx
    except dm.DepthMatchError as e:
        ok = "unsupported" in str(e).lower() or "ratio" in str(e).lower()
        print("   (rejected: %s)" % e)
    report("multiscale", "ratios=%s win=%d %dx%d C=%d" % (ratios, maxh, H, W, C), ok)

    # ---- stand-alone extractOutput on random probabilities
    h, w, n = int(rng.integers(1, 20)), int(rng.integers(1, 40)), int(rng.choice([1, 3, 9, 33, 64, 289]))
    p = (rng.random((h, w, n)) * float(rng.choice([0.05, 0.2, 0.6]))).astype(np.float32)
    thr = float(rng.choice([0.11, 0.21, 0.0, 0.5]))
    r0, s0 = rng.integers(-3, 3, (h, w)).astype(np.int64), rng.random((h, w)).astype(np.float32)
    ret, sc, _ = O.extract_output(p, thr, r0.copy(), s0.copy())
    gr, gs = r0.copy(), s0.copy()
    dm.extractoutput.extractOutput(p, gs, thr, gr)
    report("extract", "%dx%dx%d thr=%.2f" % (h, w, n, thr), np.array_equal(gr, ret) and np.array_equal(gs, sc))

    # ---- post-processing
    h, w, k = int(rng.integers(6, 40)), int(rng.integers(6, 60)), int(rng.choice([1, 2, 3, 4, 5]))
    flow = np.clip(np.rint(rng.normal(0, 3, (2, h, w))), -7, 8).astype(np.float32) + \
        (rng.random((2, h, w)).astype(np.float32) - 0.5) * 0.8
    mask = (rng.random((h, w)) > rng.random()).astype(np.float32)
    ok = True
    for method in ("med", "max"):
        ok &= np.array_equal(dm.postProcessImage(flow, mask, k, method), O.post_process_image(flow, mask, k, method))
    ix, iy = int(rng.integers(0, 9)), int(rng.integers(0, 9))
    ok &= np.array_equal(dm.enlargeMask(mask.copy(), ix, iy), O.enlarge_mask(mask, ix, iy))
    report("postprocess", "%dx%d k=%d enlarge=(%d,%d)" % (h, w, k, ix, iy), bool(ok))

    # ---- warps
    c, hs, ws = int(rng.integers(1, 5)), int(rng.integers(2, 30)), int(rng.integers(2, 40))
    src = rng.random((c, hs, ws)).astype(np.float32)
    hd, wd = int(rng.integers(1, 30)), int(rng.integers(1, 40))
    field = np.stack([rng.uniform(-3, hs + 2, (hd, wd)), rng.uniform(-3, ws + 2, (hd, wd))]).astype(np.float32)
    ok = np.array_equal(dm.cartesian2polar(src, field), O.warp_bilinear(src, field))
    a = rng.uniform(-0.1, 0.1)
    Hm = np.array([[np.cos(a), -np.sin(a), rng.uniform(-3, 3)], [np.sin(a), np.cos(a), rng.uniform(-3, 3)],
                   [rng.uniform(-1e-3, 1e-3), rng.uniform(-1e-3, 1e-3), 1.0]])
    wo, mo = O.warp_homography(src, Hm)
    wg, mg = dm.warpHomography(src, Hm)
    ok &= np.array_equal(mg, mo) and np.allclose(wg, wo, rtol=0, atol=1e-6)
    report("warps", "%dx%dx%d -> %dx%d" % (c, hs, ws, hd, wd), bool(ok))
print("failures:", bad)
sys.exit(1 if bad else 0)
