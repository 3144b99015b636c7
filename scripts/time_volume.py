"""Kernel times of volume mode at north size (one pair, 888 MB of volume): strip kernel (default)
against the tiled kernel (volume_kernel = 1), SSD and soft-max volume, with and without the stores
(volume_debug = 1: tuning only).  python scripts/time_volume.py [pairs] [volume_debug values, e.g. 0,1]
DM_ROOT=<other checkout> runs another build of the library for A/B comparisons on one box."""
import os, sys
ROOT = os.environ.get("DM_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import torch
import depthmatch as dm
g = torch.Generator(device="cuda").manual_seed(1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
f2 = torch.randn((B, 10, 360, 640), device="cuda", generator=g)
in1 = f2[:, :, 12:12 + 328, 20:20 + 608] + 0.05 * torch.randn((B, 10, 328, 608), device="cuda", generator=g)
ctx = dm.Context(0)
ctx.set_profiling(True)
nbytes = B * 328 * 608 * 1089 * 4
DBG = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1]
for kern in (0, 1, 2):
    for dbg in DBG:
        ctx.set_option("ssd_form", "diff" if kern == 2 else "auto")
        try:
            ctx.set_option("volume_kernel", kern & 1)
        except Exception:
            pass   # a build of the library older than the strip kernel (A/B runs)
        ctx.set_option("volume_debug", dbg)
        for softmax in (False, True):
            ts = []
            for _ in range(6):
                out = dm.match_volume(in1, f2, 33, 33, softmax=softmax, ctx=ctx)
                ts.append(ctx.last_kernel_ms())
            t = min(ts[2:])
            import time
            torch.cuda.synchronize(); w0 = time.perf_counter()
            for _ in range(10):
                out = dm.match_volume(in1, f2, 33, 33, softmax=softmax, ctx=ctx)
            torch.cuda.synchronize(); wall = (time.perf_counter() - w0) / 10 * 1e3
            print("kernel=%s stores=%s %-8s volume sweep ms: %s  -> %.2f TB/s" % (("strip", "tiled", "strip, difference form")[kern], ("on" if dbg == 0 else "off"),
                  "softmax" if softmax else "ssd", " ".join("%.3f" % x for x in ts[2:]), nbytes / t / 1e9), "| whole call incl. allocation of the result: %.3f ms" % wall, flush=True)
            del out
ctx.set_option("volume_debug", 0)
ctx.set_option("volume_kernel", 0)
