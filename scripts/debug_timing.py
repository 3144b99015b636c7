import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import depthmatch as dm
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C, H, W, MH, MW = 10, 360, 640, 33, 33
f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
f1[:, :, 16:16+328, 16:16+608] = f2[:, :, 20:20+328, 12:12+608] + 0.05 * torch.randn(B, C, 328, 608, device="cuda")
in1 = f1[:, :, 16:16+328, 16:16+608]
ctx = dm.Context(0); ctx.set_profiling(True)
for want in (("index",), ("index", "pmax"), ("index", "pmax", "score_thr"), ("index", "pmax", "score_thr", "soft_yx")):
    for canvas in (None, (H, W)):
        for it in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = dm.match_extract(in1, f2, MH, MW, canvas=canvas, want=want, ctx=ctx)
            t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(want, canvas, "host %.2f ms  total %.2f ms  kernel %.2f ms launches %d" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3, ctx.last_kernel_ms(), ctx.launch_count()), flush=True)
