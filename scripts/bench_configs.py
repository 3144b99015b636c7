"""Side measurements of the other BASELINE.json configurations (not the bench.py headline):
kernel-level timings with CUDA events, inputs resident in HBM."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import torch
import depthmatch as dm


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    g = torch.Generator(device="cuda").manual_seed(1)
    res = {}
    # c1-like: 320x180 -> 2-layer filter output 10x161x301, window 17x17
    in2 = torch.randn((1, 10, 161, 301), device="cuda", generator=g)
    in1 = in2[:, :, 8:8 + 145, 8:8 + 285] + 0.05 * torch.randn((1, 10, 145, 285), device="cuda", generator=g)
    ms = timed(lambda: dm.match_extract(in1, in2, 17, 17, want=("index", "pmax", "score_thr")))
    res["c1 10x161x301 17x17 (1 pair)"] = {"ms": ms, "pairs_per_s": 1e3 / ms}
    # c1 from raw frames: 3x180x320 RGB pair -> {3,5,5,8} tanh {4,16,16,10} -> 10x161x301 -> 17x17
    import numpy as np
    geo1 = dm.Geometry(layers=[[3, 5, 5, 8], [4, 16, 16, 10]], maxh=17, maxw=17, hImg=180, wImg=320)
    flt = dm.getFilter(geo1, np.random.default_rng(1))
    fr = torch.rand((2, 3, 180, 320), device="cuda", generator=g)
    model1 = dm.getModel(geo1, True, False, fused=True, filter=flt)
    ms = timed(lambda: model1.forward(dm.prepareInput(geo1, fr[0], fr[1])))
    res["c1 raw frames 3x180x320: filter + 17x17 match (1 pair)"] = {"ms": ms, "pairs_per_s": 1e3 / ms}
    ms = timed(lambda: flt.forward(fr))
    res["c1 filter alone (both frames)"] = {"ms": ms}
    # the robot-host function (depth_estimation_api.lua:nextFrameDepth) on device-resident 320x180 frames
    geoA = dm.Geometry(layers=[[3, 5, 5, 8], [4, 16, 16, 10]], maxh=17, maxw=17, hImg=180, wImg=320)
    apiA = dm.DepthEstimationAPI(geoA, dm.getFilter(geoA, np.random.default_rng(2)),
                                 K=np.array([[293.8, 0, 310.4], [0, 300.6, 251.6], [0, 0, 1.0]]), first_frame=fr[0])
    Rm = np.eye(3)
    ms = timed(lambda: apiA.nextFrameDepth(fr[1], R=Rm, nFound=100, nInliers=90))
    res["nextFrameDepth 3x180x320 (warp, filter, fused match + mean extraction, masks)"] = {"ms": ms, "frames_per_s": 1e3 / ms}
    # c2: 64 pairs of 320x180, 33x33
    in2 = torch.randn((64, 10, 180, 320), device="cuda", generator=g)
    in1 = in2[:, :, 16:16 + 148, 16:16 + 288] + 0.05 * torch.randn((64, 10, 148, 288), device="cuda", generator=g)
    ms = timed(lambda: dm.match_extract(in1, in2, 33, 33, want=("index", "pmax", "score_thr")))
    res["c2 64 x 320x180 33x33"] = {"ms": ms, "pairs_per_s": 64e3 / ms}
    # c3: multiscale 640x360, ratios 1,2,4, 8x8
    geo = dm.Geometry(maxh=8, maxw=8, ratios=[1, 2, 4], multiscale=True, hImg=360, wImg=640, output_extraction_method="max")
    inp = []
    for r in (1, 2, 4):
        h, w = 360 // r, 640 // r
        f2 = torch.randn((10, h + 7, w + 7), device="cuda", generator=g)
        inp.append((f2[:, 3:3 + h, 3:3 + w].contiguous(), f2))
    model = dm.getModelMultiscale(geo, True, True)
    ms = timed(lambda: model.forward(inp))
    res["c3 multiscale 640x360 {1,2,4} 8x8"] = {"ms": ms, "pairs_per_s": 1e3 / ms}
    # c3 from raw frames: r x r average, zero padding, shared {3,5,5,10} filter per scale, then the model
    geo3 = dm.Geometry(maxh=8, maxw=8, ratios=[1, 2, 4], multiscale=True, hImg=360, wImg=640, wPatch2=5, hPatch2=5,
                       layers=[[3, 5, 5, 10]], share_filters=True, output_extraction_method="max")
    flt3 = dm.getFilter(geo3, np.random.default_rng(3))
    fr3 = torch.rand((2, 3, 360, 640), device="cuda", generator=g)
    ms = timed(lambda: model.forward(dm.multiscaleInputs(geo3, flt3, fr3[0], fr3[1])))
    res["c3 from raw frames: 3 x (average, pad, filter) per frame + multiscale model"] = {"ms": ms, "pairs_per_s": 1e3 / ms}
    # c4 from raw frames: polar remap of both frames, 1x17 / 17x1 filter, radial matcher, unmap, depth
    netp = dict(wImg=640, hImg=360, wInput=400, hInput=400, wKernel=17, hKernel=17, hWin=15,
                layers=[[3, 1, 17, 5], [5, 17, 1, 10]])
    tester = dm.RadialTester(netp, dm.getRadialFilter(netp, np.random.default_rng(4)))
    ms = timed(lambda: tester.forward(fr3[0], fr3[1], (320.73, 172.48)))
    res["c4 from raw frames: polar remap x2, filter x2, radial match, unmap, flow2depth"] = {"ms": ms, "pairs_per_s": 1e3 / ms}
    # c4: polar remap 640x360 -> 400x(400+16), radial matcher hWin 15 on 10x384x400
    img = torch.rand((3, 360, 640), device="cuda", generator=g)
    e2 = (320.73, 172.48)
    rmax = dm.getRMax(360, 640, e2)
    ms_remap = timed(lambda: dm.cartesian2polar(img, wdst=400, hdst=400, xcenter=e2[0], ycenter=e2[1], lpadding=8, rpadding=8, rmax=rmax))
    f2 = torch.randn((10, 384, 400), device="cuda", generator=g)
    f1 = f2[:, 5:5 + 370].contiguous()
    rm = dm.nn.SpatialRadialMatching(15)
    ms_match = timed(lambda: rm.argmin_flow([f1, f2]))
    res["c4 polar remap 3x360x640 -> 400x416"] = {"ms": ms_remap}
    res["c4 radial match 10x384x400 hWin 15"] = {"ms": ms_match}
    # north, flow only (winner-take-all epilogue: no soft-max)
    in2n = torch.randn((16, 10, 360, 640), device="cuda", generator=g)
    in1n = in2n[:, :, 16:16 + 328, 16:16 + 608] + 0.05 * torch.randn((16, 10, 328, 608), device="cuda", generator=g)
    outn = {"index": torch.empty((16, 328, 608), dtype=torch.int64, device="cuda"),
            "flow_full": torch.empty((16, 2, 360, 640), device="cuda")}
    ms = timed(lambda: dm.match_extract(in1n, in2n, 33, 33, canvas=(360, 640), want=("index",), out=outn))
    res["north 16 pairs, flow only (index + canvas, no scores)"] = {"ms": ms, "pairs_per_s": 16e3 / ms}
    del in2n, in1n, outn
    # c5: 1080p, 65x65, one pair and one of 8 row bands
    in2 = torch.randn((10, 1080, 1920), device="cuda", generator=g)
    in1 = in2[:, 32:32 + 1016, 32:32 + 1856] + 0.05 * torch.randn((10, 1016, 1856), device="cuda", generator=g)
    ms = timed(lambda: dm.match_extract(in1, in2, 65, 65, want=("index", "pmax", "score_thr")), reps=3, warm=1)
    res["c5 1920x1080 65x65 (1 pair, 1 GPU)"] = {"ms": ms, "pairs_per_s": 1e3 / ms, "alu_frac": 2 * 10 * 4225 * 1016 * 1856 / (ms / 1e3) / (148 * 128 * 1.965e9)}
    a, b = in1[:, :127], in2[:, :127 + 64]
    ms = timed(lambda: dm.match_extract(a, b, 65, 65, want=("index", "pmax", "score_thr")), reps=3, warm=1)
    res["c5 one of 8 row bands (127 rows + 64 halo)"] = {"ms": ms}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
