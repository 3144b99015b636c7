import sys, time, os, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import depthmatch as dm
B = 16
C, H, W, MH, MW = 10, 360, 640, 33, 33
f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
f1[:, :, 16:16+328, 16:16+608] = f2[:, :, 20:20+328, 12:12+608] + 0.05 * torch.randn(B, C, 328, 608, device="cuda")
in1 = f1[:, :, 16:16+328, 16:16+608]
ctx = dm.Context(0)
want = ("index", "pmax", "score_thr")
def run(n, label, prealloc=False):
    outs = None
    if prealloc:
        outs = {"index": torch.empty((B, 328, 608), dtype=torch.int64, device="cuda"),
                "pmax": torch.empty((B, 328, 608), device="cuda"), "score_thr": torch.empty((B, 328, 608), device="cuda"),
                "flow_full": torch.empty((B, 2, H, W), device="cuda")}
    for _ in range(3): dm.match_extract(in1, f2, MH, MW, canvas=(H, W), want=want, ctx=ctx, out=outs)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    t0 = time.perf_counter()
    evs[0].record()
    for i in range(n):
        dm.match_extract(in1, f2, MH, MW, canvas=(H, W), want=want, ctx=ctx, out=outs)
        evs[i + 1].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    print(label, "host enqueue %.1f ms; per-step ms:" % ((t1 - t0) * 1e3), " ".join("%.1f" % m for m in ms), flush=True)
run(20, "plain      ")
run(20, "prealloc   ", True)
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = False
def poll():
    while not stop:
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h); time.sleep(0.02)
t = threading.Thread(target=poll, daemon=True); t.start(); time.sleep(0.3)
run(20, "nvml poll  ")
run(20, "nvml+preall", True)
stop = True
