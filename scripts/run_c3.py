import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import torch, depthmatch as dm
g = torch.Generator(device="cuda").manual_seed(1)
geo = dm.Geometry(maxh=8, maxw=8, ratios=[1, 2, 4], multiscale=True, hImg=360, wImg=640, output_extraction_method="max")
inp = []
for r in (1, 2, 4):
    h, w = 360 // r, 640 // r
    f2 = torch.randn((10, h + 7, w + 7), device="cuda", generator=g)
    inp.append((f2[:, 3:3 + h, 3:3 + w].contiguous(), f2))
model = dm.getModelMultiscale(geo, True, True)
import time
for i in range(4):
    torch.cuda.synchronize(); t = time.perf_counter()
    out = model.forward(inp)
    torch.cuda.synchronize(); print("c3 forward ms", (time.perf_counter() - t) * 1e3)
