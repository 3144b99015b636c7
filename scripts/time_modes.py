"""Kernel times (CUDA events inside the library) of the fused sweep's variants at north size, 16 pairs:
options sweep = 0 (one-row, default) / 2 (two-row, unrolled rows) / 3 (two-row, rolled), with and
without the epilogue (volume_debug = 1: tuning only, results are wrong)."""
import os, sys
ROOT = os.environ.get("DM_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import torch
import depthmatch as dm
g = torch.Generator(device="cuda").manual_seed(1)
B = 16
f2 = torch.randn((B, 10, 360, 640), device="cuda", generator=g)
in1 = f2[:, :, 12:12 + 328, 20:20 + 608] + 0.05 * torch.randn((B, 10, 328, 608), device="cuda", generator=g)
ctx = dm.Context(0)
ctx.set_profiling(True)
for sweep in ((0,) if os.environ.get("DM_ROOT") or "--default" in sys.argv else (0, 3, 2)):
    for dbg in (0, 4, 5, 7):
        if sweep == 0 and dbg:
            continue
        ctx.set_option("sweep", sweep)
        ctx.set_option("volume_debug", dbg)
        for name, want in (("scores", ("index", "pmax", "score_thr")), ("wta", ("index",))):
            ts = []
            for _ in range(6):
                dm.match_extract(in1, f2, 33, 33, want=want, canvas=(360, 640), ctx=ctx)
                ts.append(ctx.last_kernel_ms())
            print("sweep=%d dbg=%d %-6s kernel ms: %s" % (sweep, dbg, name, " ".join("%.3f" % t for t in ts[2:])), flush=True)
ctx.set_option("volume_debug", 0)
