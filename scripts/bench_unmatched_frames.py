"""Data dependence of the fused path: four pairs of unrelated noise frames (no match inside the
window, flat soft-max) -- the thresholded output then sends more than half of the pixels to the
exact per-pixel pass.  Prints step and sweep times per output set and the exact-pass count."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import torch
import depthmatch as dm
B = 4
g = torch.Generator(device="cuda").manual_seed(3)
f1 = torch.randn((B, 10, 360, 640), device="cuda", generator=g)
f2 = torch.randn((B, 10, 360, 640), device="cuda", generator=g)
in1 = f1[:, :, 16:16 + 328, 16:16 + 608]
ctx = dm.Context(0); ctx.set_profiling(True)
def t(want):
    for _ in range(2): dm.match_extract(in1, f2, 33, 33, want=want, ctx=ctx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dm.match_extract(in1, f2, 33, 33, want=want, ctx=ctx)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5, ctx.last_kernel_ms()
for want in (("index",), ("index", "pmax"), ("index", "pmax", "score_thr")):
    print(want, "step %.3f ms sweep %.3f ms (4 unrelated noise pairs)" % t(want))
os.environ["DM_DEBUG_TODO"] = "1"
dm.match_extract(in1, f2, 33, 33, want=("index", "pmax", "score_thr"), ctx=ctx)
