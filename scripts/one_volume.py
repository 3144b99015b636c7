"""One soft-max volume call at north size (for ncu): python scripts/one_volume.py [volume_kernel] [ssd_form]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import torch
import depthmatch as dm
g = torch.Generator(device="cuda").manual_seed(1)
f2 = torch.randn((1, 10, 360, 640), device="cuda", generator=g)
in1 = f2[:, :, 12:12 + 328, 20:20 + 608] + 0.05 * torch.randn((1, 10, 328, 608), device="cuda", generator=g)
ctx = dm.Context(0)
ctx.set_option("volume_kernel", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ctx.set_option("ssd_form", sys.argv[2] if len(sys.argv) > 2 else "auto")
for softmax in (False, True, False, True):
    out = dm.match_volume(in1, f2, 33, 33, softmax=softmax, ctx=ctx)
    torch.cuda.synchronize()
    del out
