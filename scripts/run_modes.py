"""One launch of every kernel family at its benchmark size, for ncu (the second round is the one to
profile): the fused sweep (scores, flow only, soft mean) + exact thresholded pass + rescore at north
size (4 pairs), the volume kernels (2 pairs), the multiscale ring join (c3), the radial matcher
(c4), the feature extractor (c1's filter, CUDA-core kernel and the opt-in tcgen05 kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import numpy as np
import torch
import depthmatch as dm
g = torch.Generator(device="cuda").manual_seed(1)
f2 = torch.randn((4, 10, 360, 640), device="cuda", generator=g)
in1 = f2[:, :, 12:12 + 328, 20:20 + 608] + 0.05 * torch.randn((4, 10, 328, 608), device="cuda", generator=g)
ctx = dm.Context(0)
# c3: three scales {1,2,4}, 8x8 windows, prefiltered maps
geo = dm.Geometry(maxh=8, maxw=8, ratios=[1, 2, 4], multiscale=True, hImg=360, wImg=640, output_extraction_method="max")
pyr = []
for r in (1, 2, 4):
    h, w = 360 // r, 640 // r
    b = torch.randn((10, h + 7, w + 7), device="cuda", generator=g)
    pyr.append((b[:, 3:3 + h, 3:3 + w].clone(), b))
ms = dm.getModelMultiscale(geo, True, True)
# c4: radial matcher on 400-wide polar maps, hWin 15
rf2 = torch.randn((10, 384, 400), device="cuda", generator=g)
rf1 = rf2[:, 5:5 + 370].clone()
rad = dm.nn.SpatialRadialMatching(15)
flt = dm.getFilter(dm.Geometry(layers=[[3, 5, 5, 8], [4, 16, 16, 10]]), np.random.default_rng(0), ctx=ctx)
frames = torch.rand((2, 3, 360, 640), device="cuda", generator=g)
for rep in range(2):
    dm.match_extract(in1, f2, 33, 33, want=("index", "pmax", "score_thr"), canvas=(360, 640), ctx=ctx)
    dm.match_extract(in1, f2, 33, 33, want=("index",), canvas=(360, 640), ctx=ctx)
    dm.match_extract(in1, f2, 33, 33, want=("soft_yx", "conf_marginal"), ctx=ctx)
    dm.match_volume(in1[:2], f2[:2], 33, 33, ctx=ctx)
    dm.match_volume(in1[:2], f2[:2], 33, 33, softmax=True, ctx=ctx)
    ms.forward(pyr)
    rad.argmin_flow([rf1, rf2])
    ctx.set_option("conv", 0)
    flt.forward(frames)
    ctx.set_option("conv", 2)
    flt.forward(frames)
    ctx.set_option("conv", 0)
    torch.cuda.synchronize()
print("ok")
