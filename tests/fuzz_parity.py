"""Randomised parity sweep of the fused kernel against the oracle (run on a GPU box):
random channel counts, window shapes, frame sizes, SSD forms and output sets."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import depthmatch as dm
import oracle_lib as O
from synth import make_pair

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
bad_total = 0
for case in range(n_cases):
    C = int(rng.choice([1, 2, 3, 4, 5, 8, 10, 11, 16, 17, 20]))
    maxh, maxw = int(rng.integers(1, 20)), int(rng.integers(1, 20))
    wide = rng.random() < 0.25   # several 128-column tiles and 15-row tile borders
    H2 = int(rng.integers(maxh + 1, maxh + (70 if wide else 40)))
    W2 = int(rng.integers(maxw + 1, maxw + (420 if wide else 150)))
    form = rng.choice(["diff", "dot", "exact"])
    noise = float(rng.choice([0.0, 0.05, 0.5]))
    in1, in2, _ = make_pair(C, H2, W2, maxh, maxw, seed=int(rng.integers(1 << 30)), noise=noise)
    if rng.random() < 0.3:   # a flat region
        in1[:, : in1.shape[1] // 2] = 0.5
        in2[:, : in2.shape[1] // 2] = 0.5
    K = maxh * maxw
    ctx0 = dm.default_context()
    # small inputs take 5-row tiles by default: keep the 15-row kernels covered
    no_small = rng.random() < 0.5
    ctx0.set_option("no_small_tiles", "1" if no_small else "0")
    ctx0.set_option("ssd_form", form if form in ("diff", "dot") else "auto")
    want = ("index", "min_ssd", "pmax", "index_thr", "score_thr", "soft_yx")
    got = dm.match_extract(in1, in2, maxh, maxw, want=want, exact=(form == "exact"))
    wta = dm.match_extract(in1, in2, maxh, maxw, want=("index", "min_ssd"), exact=(form == "exact"))
    # scores without the soft mean: in the dot form this is the two-rows-per-warp sweep
    sco = dm.match_extract(in1, in2, maxh, maxw, want=("index", "min_ssd", "pmax", "index_thr", "score_thr"),
                           exact=(form == "exact"))
    if rng.random() < 0.3:
        # the same pair inside a batch of host buffers (pipelined chunks) and as a strided crop view
        n = int(rng.integers(2, 7))
        pos = int(rng.integers(n))
        b2 = rng.standard_normal((n,) + in2.shape).astype(np.float32)
        b2[pos] = in2
        big1 = rng.standard_normal((n, C, in1.shape[1] + 3, in1.shape[2] + 5)).astype(np.float32)
        big1[pos, :, 1:1 + in1.shape[1], 2:2 + in1.shape[2]] = in1
        view = big1[:, :, 1:1 + in1.shape[1], 2:2 + in1.shape[2]]
        batch = dm.match_extract(view, b2, maxh, maxw, want=want, exact=(form == "exact"))
        for name in want:
            if not np.array_equal(batch[name][pos], got[name]):
                print("   batch/view mismatch in", name)
                got = {k: v[pos] for k, v in batch.items()}
                got["index"] = got["index"] * 0 - 1
                break
    vol = O.spatial_matching(in1, in2, maxh, maxw).reshape(-1, K)
    prob = O.neg_softmax(vol)
    middle = (math.ceil(maxh / 2) - 1) * maxw + math.ceil(maxw / 2)
    idx, pmax = O.argmax_tie(prob, K, middle)
    gap = O.top2_relgap(prob, K) if K > 1 else np.ones(len(idx), np.float32)
    tie = gap < 1e-5
    errs = []
    if ((got["index"].reshape(-1) != idx) & ~tie).any():
        errs.append("index")
    if ((wta["index"].reshape(-1) != idx) & ~tie).any():
        errs.append("wta index")
    if not np.allclose(got["pmax"].reshape(-1), pmax, rtol=1e-4):
        errs.append("pmax")
    mn = vol.min(-1)
    tol = dict(rtol=0, atol=0) if form == "exact" else dict(rtol=1e-5, atol=1e-6)
    if not np.allclose(got["min_ssd"].reshape(-1)[~tie], mn[~tie], **tol):
        errs.append("min_ssd")
    ym, xm = O.soft_mean(prob, maxh, maxw)
    if not np.allclose(got["soft_yx"][0].reshape(-1), ym, rtol=1e-4, atol=1e-4):
        errs.append("soft")
    h1, w1 = in1.shape[1:]
    ret, sc, _ = O.extract_output(prob.reshape(h1, w1, K), 0.11, np.zeros((h1, w1), np.int64), np.zeros((h1, w1), np.float32))
    near = (np.abs(prob - 0.11) < 2e-4).any(-1).reshape(h1, w1)
    tie2 = tie.reshape(h1, w1)   # the two largest probabilities within 1e-5: either may lead the list
    if ((got["index_thr"] != ret) & ~near & ~tie2).any() or \
            not np.allclose(got["score_thr"][~near], sc[~near], rtol=1e-4, atol=1e-6):
        errs.append("thr")
        if len(sys.argv) > 3:   # verbose: where and what
            badpx = np.argwhere((((got["index_thr"] != ret) & ~tie2) | ~np.isclose(got["score_thr"], sc, rtol=1e-4, atol=1e-6)) & ~near)
            print("   small tiles off:", no_small, "bad pixels", badpx[:6].tolist())
            for (yy, xx) in badpx[:3]:
                pr = prob.reshape(h1, w1, K)[yy, xx]
                print("   px", yy, xx, "got", got["index_thr"][yy, xx], got["score_thr"][yy, xx], "want", ret[yy, xx], sc[yy, xx],
                      "probs>0.1:", [(int(k) + 1, float(pr[k])) for k in np.nonzero(pr > 0.1)[0]])
    if ((sco["index"].reshape(-1) != idx) & ~tie).any():
        errs.append("scores index")
    if not np.allclose(sco["pmax"].reshape(-1), pmax, rtol=1e-4):
        errs.append("scores pmax")
    if not np.allclose(sco["min_ssd"].reshape(-1)[~tie], mn[~tie], **tol) or \
            not np.allclose(wta["min_ssd"].reshape(-1)[~tie], mn[~tie], **tol):
        errs.append("scores/wta min_ssd")
    if ((sco["index_thr"] != ret) & ~near & ~tie2).any() or \
            not np.allclose(sco["score_thr"][~near], sc[~near], rtol=1e-4, atol=1e-6):
        errs.append("scores thr")
    status = "ok" if not errs else "FAIL " + ",".join(errs)
    bad_total += bool(errs)
    print("case %2d C=%2d win=%2dx%2d in2=%3dx%3d form=%-5s noise=%.2f ties=%d: %s"
          % (case, C, maxh, maxw, H2, W2, form, noise, int(tie.sum()), status), flush=True)
print("failures:", bad_total)
sys.exit(1 if bad_total else 0)
