"""Generates tests/golden/*.npz.  Run in the authoring container (where /root/reference is
mounted and oracle/_ref is built from the reference's own sources):

    python tests/golden/make_golden.py

Two kinds of vectors:
  * ref_*.npz    -- inputs + outputs of the REFERENCE's own native code (oracle/_ref:
                    version2/extract_output.cpp, x2yxMulti2.c and the inline.load C bodies of
                    the Lua files, compiled as-is).  These pin the
                    oracle (tests/test_golden.py, CPU) and the CUDA kernels (GPU).
  * oracle_*.npz -- inputs + outputs of the oracle restatement for the parts whose arithmetic
                    lives in un-vendored Torch7 code (matching, softmax, cascade, polar remap):
                    regression vectors, they pin the CUDA path to the oracle across rounds and
                    the oracle to itself (parity with the real reference is unpinned there).
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
from synth import make_pair  # noqa: E402


def inline_vectors():
    """ref_inline.npz: the reference's inline.load C bodies (oracle/extract_inline.py compiles them
    verbatim into oracle/_ref): postProcessImage fmed/fmax, enlargeMask, radial(), polar LUTs,
    flow2depth."""
    rng = np.random.default_rng(20261019)
    h, w = 36, 64
    base = np.clip(np.rint(rng.normal(0, 3, (2, h, w))), -7, 8).astype(np.float32)
    flow_med = base + rng.random((2, h, w)).astype(np.float32)
    flow_max = base + (rng.random((2, h, w)).astype(np.float32) - 0.5) * 0.8
    mask = (rng.random((h, w)) > 0.3).astype(np.float32)
    mask[10:16, 20:30] = 0
    out = {"flow_med": flow_med, "flow_max": flow_max, "mask": mask}
    for k in (3, 5):
        out["med%d" % k] = O.post_process_image(flow_med, mask, k, "med", "ref")
        out["max%d" % k] = O.post_process_image(flow_max, mask, k, "max", "ref")
    emask = (rng.random((h, w)) > 0.35).astype(np.float32)
    emask[5] = 0
    emask[:, 9] = 0
    out["emask"] = emask
    out["emask_4_3"] = O.enlarge_mask(emask, 4, 3, "ref")
    out["emask_16_16"] = O.enlarge_mask(emask, 16, 16, "ref")
    rflow = (rng.normal(0, 2, (2, h, w)) * (rng.random((2, h, w)) > 0.2)).astype(np.float32)
    rd, rc = O.radial_depth(rflow, 25.3, 31.7, w / 2, "ref")
    out.update(rflow=rflow, rcentre=np.array([25.3, 31.7, w / 2]), rdepth=rd, rconf=rc)
    e2 = (641.4552 * 80 / 1280.0, 344.950836 * 80 / 1280.0)
    rmax = O.get_rmax(45, 80, *e2)
    out.update(polar_geom=np.array([45, 80, 50, 48]), polar_e2=np.array(e2), polar_rmax=np.float64(rmax),
               c2p=O.ref_c2p_mask(48, 50, e2[0], e2[1], rmax), p2c=O.ref_p2c_mask(48, 50, 80, 45, e2[0], e2[1], rmax))
    pflow = (rng.random((50, 48)) * 9 * (rng.random((50, 48)) > 0.1)).astype(np.float32)
    fd, fc = O.ref_flow2depth(pflow, e2[0], e2[1], 1000.0)
    out.update(pflow=pflow, fdepth=fd, fconf=fc)
    np.savez_compressed(os.path.join(HERE, "ref_inline.npz"), **out)


def calibration_vectors():
    """ref_calibration.npz: the four Torch7-serialised calibration tables the reference ships
    (radial/*.cal, version2/rectified_gopro.cal; written by radial/generate_calibration_file.lua),
    byte for byte -- fixtures for the torch7io reader/writer."""
    ref = "/root/reference"
    out = {}
    for name, rel in (("ardrone", "radial/ardrone.cal"), ("gopro", "radial/gopro.cal"),
                      ("rectified_gopro", "radial/rectified_gopro.cal"),
                      ("rectified_gopro_v2", "version2/rectified_gopro.cal")):
        out[name] = np.frombuffer(open(os.path.join(ref, rel), "rb").read(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_calibration.npz"), **out)


def c1_cars_vectors():
    """oracle_c1_cars.npz: BASELINE config 1 -- the reference's own test pair celiu/car1.jpg,
    car2.jpg (640x480), scaled to 320x180 like opticalflow.lua:139-140, through the default
    two-layer filter {3,5,5,8} tanh {4,16,16,10} with seeded random-init weights and a 17x17
    window.  Frames are stored as uint8 RGB at 320x180; features, volume statistics and indices
    come from the oracle (Torch7 is not available to run the reference itself)."""
    import cv2
    rng = np.random.default_rng(1)
    frames = []
    for name in ("car1.jpg", "car2.jpg"):
        bgr = cv2.imread(os.path.join("/root/reference/celiu", name), cv2.IMREAD_COLOR)
        assert bgr is not None and bgr.shape == (480, 640, 3), name
        small = cv2.resize(bgr, (320, 180), interpolation=cv2.INTER_AREA)
        frames.append(np.ascontiguousarray(small[:, :, ::-1].transpose(2, 0, 1)))      # RGB, CHW, uint8
    frames = np.stack(frames)
    stdv1 = 1.0 / math.sqrt(5 * 5 * 3)
    w1 = rng.uniform(-stdv1, stdv1, (8, 3, 5, 5)).astype(np.float32)
    b1 = rng.uniform(-stdv1, stdv1, 8).astype(np.float32)
    conn = np.array([[f + 1, o + 1] for o in range(10) for f in sorted(rng.permutation(8)[:4])], np.int32)
    stdv2 = 1.0 / math.sqrt(16 * 16 * 4)
    w2 = rng.uniform(-stdv2, stdv2, (40, 16, 16)).astype(np.float32)
    b2 = rng.uniform(-stdv2, stdv2, 10).astype(np.float32)
    layers = [dict(weight=w1, bias=b1, tanh=True), dict(weight=w2, bias=b2, conn=conn)]
    img = frames.astype(np.float32) / 255.0
    f1 = O.filter_forward(img[0], layers)                         # 10 x 161 x 301
    f2 = O.filter_forward(img[1], layers)
    maxh = maxw = 17
    in1 = np.ascontiguousarray(f1[:, 8:8 + 161 - 16, 8:8 + 301 - 16])   # prepareInput's crop
    vol = O.spatial_matching(in1, f2, maxh, maxw)
    prob = O.neg_softmax(vol)
    K = maxh * maxw
    idx, pmax = O.argmax_tie(prob, K, 8 * 17 + 9)
    np.savez_compressed(os.path.join(HERE, "oracle_c1_cars.npz"), frames=frames, w1=w1, b1=b1, conn=conn, w2=w2,
                        b2=b2, index=idx.reshape(145, 285).astype(np.int16),
                        pmax_sample=pmax.reshape(145, 285)[::4, ::4].copy(),
                        near_tie=np.packbits(O.top2_relgap(prob, K).reshape(145, 285) < 1e-5),
                        feat2_sample=f2[:, ::20, ::20].copy())


def main():
    assert O.ref() is not None, "build oracle/_ref first (make -C oracle)"
    if "--cars-only" in sys.argv:
        c1_cars_vectors()
        return
    if "--calibration-only" in sys.argv:
        calibration_vectors()
        return
    if "--inline-only" in sys.argv:
        inline_vectors()
        return
    inline_vectors()
    calibration_vectors()
    c1_cars_vectors()
    rng = np.random.default_rng(20261018)

    # ---- reference native code: extractOutput / extractOutputMarginalized
    out = {}
    for i, (n, thr) in enumerate([(9, 0.11), (33, 0.11), (289, 0.11), (64, 0.21), (25, 0.0)]):
        inp = (rng.random((5, 8, n), dtype=np.float32) * (0.45 if n < 64 else 0.2)).astype(np.float32)
        inp[0, 0] = 0
        inp[1, 1] = inp[1, 1, 0]
        r0 = rng.integers(-3, 3, (5, 8)).astype(np.int64)
        s0 = rng.random((5, 8)).astype(np.float32)
        ret, sc, _ = O.extract_output(inp, thr, r0, s0, "ref")
        mret, mgd = O.extract_output_marginalized(inp, thr, 0.6, r0, "ref")
        out.update({"in%d" % i: inp, "thr%d" % i: np.float64(thr), "r0_%d" % i: r0, "s0_%d" % i: s0,
                    "ret%d" % i: ret, "sc%d" % i: sc, "mret%d" % i: mret, "mgd%d" % i: mgd})
    np.savez_compressed(os.path.join(HERE, "ref_extract_output.npz"), **out)

    # ---- reference native code: x2yxMulti2.c
    out = {}
    for i, (mh, mw, rat) in enumerate([(8, 8, [1, 2]), (8, 8, [1, 2, 4]), (16, 16, [1, 2, 4])]):
        L = O.multiscale_length(mh, mw, rat)
        x = np.arange(-2, L + 38 + (L % 2), dtype=np.int64).reshape(2, -1)
        ry, rx = O.x2yx_multi2_bugcompat(x, mh, mw, rat, "ref", fill=0)
        out.update({"x%d" % i: x, "geom%d" % i: np.array([mh, mw] + rat), "rety%d" % i: ry, "retx%d" % i: rx})
    np.savez_compressed(os.path.join(HERE, "ref_x2yxmulti2.npz"), **out)

    # ---- oracle: single-scale fused path
    maxh, maxw = 5, 7
    in1, in2, flow = make_pair(3, 28, 46, maxh, maxw, seed=99, noise=0.4)
    K = maxh * maxw
    vol = O.spatial_matching(in1, in2, maxh, maxw)
    prob = O.neg_softmax(vol)
    middle = (math.ceil(maxh / 2) - 1) * maxw + math.ceil(maxw / 2)
    idx, pmax = O.argmax_tie(prob, K, middle)
    shp = in1.shape[1:]
    ret, sc, _ = O.extract_output(prob.reshape(shp + (K,)), 0.11)
    ym, xm = O.soft_mean(prob, maxh, maxw)
    np.savez_compressed(os.path.join(HERE, "oracle_single_scale.npz"), in1=in1, in2=in2, flow=flow,
                        window=np.array([maxh, maxw]), volume=vol, prob=prob.astype(np.float32),
                        index=idx.reshape(shp), pmax=pmax.reshape(shp), index_thr=ret, score_thr=sc,
                        soft_y=ym.reshape(shp), soft_x=xm.reshape(shp),
                        gap=O.top2_relgap(prob, K).reshape(shp),
                        canvas=O.flow_canvas(idx, shp[0], shp[1], maxh, maxw, 28, 46))

    # ---- oracle: cascade + ring join + radial
    ratios = [1, 2, 4]
    casc_in = rng.random((3, 11, 8, 8)).astype(np.float32)
    casc = O.cascade_add(casc_in, ratios)
    f2 = rng.standard_normal((4, 30, 20)).astype(np.float32)
    f1 = (np.roll(f2, -3, axis=1)[:, :22] + 0.1 * rng.standard_normal((4, 22, 20))).astype(np.float32)
    rvol = O.radial_matching(f1, f2, 9)
    ridx, rmin = O.argmin_tie(rvol, 9, 0)
    np.savez_compressed(os.path.join(HERE, "oracle_multiscale_radial.npz"), casc_in=casc_in, casc=casc,
                        ring=O.ring_join(casc, ratios), ratios=np.array(ratios), rf1=f1, rf2=f2,
                        rvol=rvol, rflow=(ridx - 1).reshape(22, 20).astype(np.float32))

    # ---- oracle: polar remap
    hImg, wImg, hIn, wIn = 45, 80, 50, 48
    e2 = (641.4552 * wImg / 1280.0, 344.950836 * wImg / 1280.0)
    img = rng.random((2, hImg, wImg)).astype(np.float32)
    rmax = O.get_rmax(hImg, wImg, *e2)
    m = O.c2p_mask(wIn, hIn, e2[0], e2[1], 2, 3, rmax, 1.0)
    pol = O.warp_bilinear(img, m)
    m2 = O.p2c_mask(wIn, hIn, wImg, hImg, e2[0], e2[1], rmax, 1.0)
    back = O.warp_bilinear(pol[:, :, 2:2 + wIn], m2)
    np.savez_compressed(os.path.join(HERE, "oracle_polar.npz"), img=img, e2=np.array(e2), rmax=np.float64(rmax),
                        geom=np.array([hImg, wImg, hIn, wIn, 2, 3]), c2p=m, polar=pol, p2c=m2, back=back)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
