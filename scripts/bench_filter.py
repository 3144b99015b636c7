"""Times the feature extractor (next row 3) on device-resident frames: c1's two-layer filter on
640x360 RGB pairs, the multiscale {3,5,5,10} prefilter, the radial net.  CUDA events on the
context's stream; prints lane-FMA/s against the measured FP32 pipe peak."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import depthmatch as dm  # noqa: E402

PEAK = 37.2e12  # lane-FMA/s, 148 SMs x 128 lanes x 1.965 GHz


def fmas(flt, h, w, n):
    total = 0
    for m in flt.modules:
        if not hasattr(m, "weight"):
            continue
        h, w = h - m.kH + 1, w - m.kW + 1
        nconn = m.connTable.shape[0] if m.connTable is not None else m.nInputPlane * m.nOutputPlane
        total += nconn * m.kH * m.kW * h * w
    return total * n


def run(name, flt, x, iters=20):
    for _ in range(3):
        out = flt.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = flt.forward(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    f = fmas(flt, x.shape[-2], x.shape[-1], x.shape[0])
    print(json.dumps({"filter": name, "in": list(x.shape), "out": list(out.shape), "ms": round(ms, 4),
                      "gfma": round(f / 1e9, 3), "alu_frac": round(f / (ms * 1e-3) / PEAK, 3)}))


def main():
    rng = np.random.default_rng(0)
    if "--conv" in sys.argv:  # 2 = the tcgen05 kernel for the layers it can plan (filter_tc.cu)
        dm.default_context().set_option("conv", int(sys.argv[sys.argv.index("--conv") + 1]))
    g = dm.Geometry(layers=[[3, 5, 5, 8], [4, 16, 16, 10]])
    x = torch.rand((2, 3, 360, 640), device="cuda")
    if "--one" in sys.argv:  # a short run for ncu
        run("same, 16 pairs", dm.getFilter(g, rng), torch.rand((32, 3, 360, 640), device="cuda"), iters=2)
        return
    run("c1 {3,5,5,8} tanh {4,16,16,10} map", dm.getFilter(g, rng), x)
    run("same, 16 pairs", dm.getFilter(g, rng), torch.rand((32, 3, 360, 640), device="cuda"))
    run("c3 {3,5,5,10}", dm.getFilter(dm.Geometry(layers=[[3, 5, 5, 10]]), rng), x)
    run("c4 radial {3,1,17,5},{5,17,1,10}", dm.getRadialFilter(dict(layers=[[3, 1, 17, 5], [5, 17, 1, 10]]), rng),
        torch.rand((2, 3, 400, 416), device="cuda"))
    run("full 10->10 17x17", dm.getRadialFilter(dict(layers=[[10, 17, 17, 10]]), rng),
        torch.rand((2, 10, 360, 640), device="cuda"))


if __name__ == "__main__":
    main()
