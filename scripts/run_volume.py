import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import torch, depthmatch as dm
B = 2
f1 = torch.randn(B, 10, 360, 640, device="cuda"); f2 = torch.randn(B, 10, 360, 640, device="cuda")
in1 = f1[:, :, 16:16+328, 16:16+608]
ctx = dm.Context(0); ctx.set_profiling(True)
for i in range(3):
    v = dm.match_volume(in1, f2, 33, 33, ctx=ctx)
    torch.cuda.synchronize()
    print("volume kernel ms", ctx.last_kernel_ms())
