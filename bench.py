#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the fused dense-matching path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA, libdepthmatch.so)
    python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle port)

Workload ("north"): frame pairs of 640x360 feature maps, C = 10 channels,
33x33 search window: in2 10x360x640, in1 = the 10x328x608 window crop of frame 1
(prepareInput, opticalflow_model.lua:131-151), synthetic (seeded N(0,1) maps, frame 1 = frame 2
displaced by a smooth integer flow + N(0, 0.05^2) noise).  One step = one call of the fused
path (dm_match_extract) over a batch of B pairs per GPU: winner index with the zero-flow tie
rule, winner probability, thresholded extractOutput score and the flow canvas.  Pairs are
independent, so N GPUs run N independent shards (weak scaling, no data-path collective); the
timed region is bracketed by barrier + synchronize and the max over ranks is taken.

One JSON line on stdout (rank 0).  `value` = pairs/s with inputs resident in HBM, timed with
CUDA events; `e2e` = the same through the C ABI with pinned HOST buffers (H2D + D2H inside the
timed region); `roofline` describes the sweep kernel; `cpu_baseline` is the oracle port timed
on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

C, H, W, MAXH, MAXW = 10, 360, 640, 33, 33
H1, W1 = H - MAXH + 1, W - MAXW + 1
K = MAXH * MAXW
# algorithmic bytes / issue slots per frame pair
BYTES_FUSED = 4 * C * (H1 * W1 + H * W) + 12 * H1 * W1
BYTES_VOLUME = BYTES_FUSED + 4 * H1 * W1 * K
ALU_SLOTS = 2 * C * K * H1 * W1
NCU_DRAM_BYTES_PER_PAIR = (75.497216e6 + 3.643392e6) / 4   # profiles/r01_ncu_fused_kernel.md, dot-form kernel
METRIC = "frame-pairs/sec @640x360, 33x33 window"
WORKLOAD = "north: 640x360 feature maps, C=10, 33x33 window, fused match+extract"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured"
    except Exception:
        return 6650.0, 1965.0, "fallback"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md), polled in-process
    through NVML (pynvml): spawning nvidia-smi in a loop takes driver locks for tens of
    milliseconds and perturbs a sub-second timed region."""

    def __init__(self, index):
        self.rows, self.ok, self._stop = [], False, False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.ok = False

    def _poll(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, [n for n, bit in names.items() if mask & bit]))
            except Exception:
                pass
            time.sleep(0.02)

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.ok and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        return len(self.rows)

    def summary(self, start, stop):
        rows = self.rows[start:stop] or self.rows[-3:]
        sm = [r[0] for r in rows]
        reasons = sorted({n for r in rows for n in r[1]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": getattr(self, "max", None),
                "reasons": reasons, "samples": len(rows)}

    def stop(self):
        self._stop = True


def make_batch(B, seed0):
    """B pairs: full frame-1 maps f1 (the crop is a view) and frame-2 maps f2."""
    from synth import make_pair
    cy, cx = (MAXH + 1) // 2 - 1, (MAXW + 1) // 2 - 1
    f1 = np.empty((B, C, H, W), np.float32)
    f2 = np.empty((B, C, H, W), np.float32)
    for b in range(B):
        in1, in2, _ = make_pair(C, H, W, MAXH, MAXW, seed=seed0 + b, noise=0.05)
        f1[b] = np.random.default_rng(seed0 + b + 7919).standard_normal((C, H, W), dtype=np.float32)
        f1[b, :, cy:cy + H1, cx:cx + W1] = in1
        f2[b] = in2
    return f1, f2


def crop(f1):
    cy, cx = (MAXH + 1) // 2 - 1, (MAXW + 1) // 2 - 1
    return f1[:, :, cy:cy + H1, cx:cx + W1]


def cpu_reference_sample(rows, nthreads, repeats=2):
    """The reference CPU path (oracle port: SpatialMatching -> Minus -> SoftMax -> argmax+tie ->
    extractOutput(0.11) -> canvas) on `rows` output rows of one north pair.  Returns s/pair."""
    import oracle_lib as O
    from synth import make_pair
    in1, in2, _ = make_pair(C, H, W, MAXH, MAXW, seed=1234, noise=0.05)
    a, b = in1[:, :rows], in2[:, :rows + MAXH - 1]
    best = None
    for it in range(repeats + 1):
        t0 = time.perf_counter()
        vol = O.spatial_matching(a, b, MAXH, MAXW, nthreads=nthreads)
        t1 = time.perf_counter()
        prob = O.neg_softmax(vol, nthreads=nthreads)
        t2 = time.perf_counter()
        idx, _ = O.argmax_tie(prob, K, (MAXH // 2) * MAXW + MAXW // 2 + 1)
        O.extract_output(prob.reshape(rows, W1, K), 0.11)
        O.flow_canvas(idx, rows, W1, MAXH, MAXW, rows + MAXH - 1, W)
        t3 = time.perf_counter()
        if it and (best is None or t3 - t0 < best[0]):
            best = (t3 - t0, t1 - t0, t2 - t1, t3 - t2)
    scale = H1 / rows
    return {"s_per_pair": best[0] * scale, "match_s": best[1] * scale, "softmax_s": best[2] * scale,
            "extract_s": best[3] * scale}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference itself needs Torch7/Lua which this image does not have), all host threads."""
    if rank != 0:
        return
    import oracle_lib as O
    O.build()
    cores = os.cpu_count() or 1
    rows = 41  # 1/8 of a pair per step
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_sample(rows, cores, repeats=1)
    t0 = time.perf_counter()
    per = [cpu_reference_sample(rows, cores, repeats=1)["s_per_pair"] for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    s_pair = float(np.mean(per))
    val = 1.0 / s_pair
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frame-pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "%d of %d output rows of one pair per step" % (rows, H1)},
            "cpu_baseline": {"value": val, "unit": "frame-pairs/s", "cores": cores, "kind": "port",
                             "sample": "%d of %d output rows of one pair, scaled; all host threads" % (rows, H1)},
            "e2e": {"value": val, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="frame pairs per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-volume", action="store_true", help="skip the volume-mode side measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import depthmatch as dm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    from depthmatch import parallel as dm_parallel
    numa_node = None if os.environ.get("DM_NO_NUMA_BIND") else dm_parallel.bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = dm.Context(local)
    ctx.set_profiling(True)
    B = args.batch
    f1_h, f2_h = make_batch(B, 1234 + rank * 1000)
    f1 = torch.from_numpy(f1_h).cuda()
    f2 = torch.from_numpy(f2_h).cuda()
    in1 = crop(f1)
    want = ("index", "pmax", "score_thr")

    # result buffers are allocated once: a cudaMalloc inside the timed region (torch's caching
    # allocator growing when two result sets are alive) would stall the device for milliseconds
    out_d = {"index": torch.empty((B, H1, W1), dtype=torch.int64, device="cuda"),
             "pmax": torch.empty((B, H1, W1), device="cuda"),
             "score_thr": torch.empty((B, H1, W1), device="cuda"),
             "flow_full": torch.empty((B, 2, H, W), device="cuda")}

    def step():
        return dm.match_extract(in1, f2, MAXH, MAXW, canvas=(H, W), want=want, ctx=ctx, out=out_d)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_first()
    # untimed settle phase: a box that just ran another process (memory scrubbing, clock ramp)
    # needs a moment before step times are steady; then the W contractual warm-up steps
    prev = None
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        e1.synchronize()
        cur = e0.elapsed_time(e1)
        if prev is not None and abs(cur - prev) < 0.03 * prev:
            break
        prev = cur
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ctx.launch_count()
    mark0 = sampler.mark() if sampler else 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    step_evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        out = step()
        step_evs[i].record()
    ev1.record()
    barrier()
    mark1 = sampler.mark() if sampler else 0
    launches = ctx.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    per_step = [(ev0 if i == 0 else step_evs[i - 1]).elapsed_time(step_evs[i]) for i in range(args.steps)]
    if os.environ.get("DM_BENCH_DEBUG"):
        print("per-step ms: " + " ".join("%.2f" % v for v in per_step), file=sys.stderr)
    t = torch.tensor([ms_total], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total / 1e3)

    # sanity: the planted flow came back (cheap, outside the timed region)
    idx0 = out["index"][0].cpu().numpy()
    assert ((idx0 >= 1) & (idx0 <= K)).all()

    # kernel-only duration: re-time a few launches one by one with the library's own events
    ks = []
    for _ in range(5):
        step()
        ks.append(ctx.last_kernel_ms())
    k_ms = float(np.mean(ks))

    # ---- e2e: pinned host buffers through the C ABI, H2D + D2H inside the timed region
    f1_p = torch.from_numpy(f1_h).pin_memory()
    f2_p = torch.from_numpy(f2_h).pin_memory()
    in1_p = crop(f1_p).numpy()
    f2_pn = f2_p.numpy()

    # results land in pinned host buffers too (what a host that consumes them every frame keeps)
    shapes = {"index": ((B, H1, W1), torch.int64), "pmax": ((B, H1, W1), torch.float32),
              "score_thr": ((B, H1, W1), torch.float32), "flow_full": ((B, 2, H, W), torch.float32)}
    out_p = {k: torch.empty(shp, dtype=dt).pin_memory().numpy() for k, (shp, dt) in shapes.items()}

    # Two contexts alternate so that the H2D of step i+1 overlaps the kernels and the D2H of step
    # i (DM_FLAG_ASYNC: the call returns once queued; a context is synchronised before its
    # buffers are reused, and both before the clock stops).  Every step copies its inputs in and
    # its results out.
    ctxs = [ctx, dm.Context(torch.cuda.current_device())]
    outs = [out_p, {k: torch.empty(shp, dtype=dt).pin_memory().numpy() for k, (shp, dt) in shapes.items()}]

    def e2e_step(i):
        c = ctxs[i % 2]
        c.synchronize()
        return dm.match_extract(in1_p, f2_pn, MAXH, MAXW, canvas=(H, W), want=want, ctx=c, out=outs[i % 2],
                                async_=True)

    for i in range(2):
        e2e_step(i)
    for c in ctxs:
        c.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.e2e_steps):
        r = e2e_step(i)
    for c in ctxs:
        c.synchronize()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # the host-buffer path returned what the device-buffer path computed (outside the timed region)
    for o in outs:
        assert np.array_equal(o["index"], out["index"].cpu().numpy()), "e2e result differs from the device path"
    te = torch.tensor([e2e_s], device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * B * args.e2e_steps / float(te.item())
    h2d = 4 * (B * C * H1 * W1 + B * C * H * W)   # the frame-1 crop is packed by a 3-D copy
    d2h = B * (H1 * W1 * (8 + 4 + 4) + 2 * H * W * 4)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    hbm_gbs, sm_max, which = peaks()
    clocks = sampler.summary(mark0, mark1) if sampler else {}
    if sampler:
        sampler.stop()
    sm_mhz = clocks.get("sm_mhz") or sm_max
    achieved = BYTES_FUSED * B / (k_ms / 1e3) / 1e9
    alu_peak = 148 * 128 * sm_max * 1e6
    line = {
        "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "mode": "fused",
                   "outputs": list(want) + ["flow_full"],
                   "l2": "inputs per step (%.0f MB) larger than the 126 MB L2" % (B * 2 * C * H * W * 4 / 1e6)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": achieved / hbm_gbs,
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel in the round's
                     # `ncu --set full` capture (4 pairs per launch: 75.5 MB + 3.6 MB), per pair x B
                     "traffic": NCU_DRAM_BYTES_PER_PAIR * B, "traffic_unit": "bytes per launch",
                     "algorithmic_bytes": BYTES_FUSED * B, "peak_source": which,
                     "kernel": "match_extract_kernel<10>", "kernel_ms": k_ms,
                     "note": "fused mode never writes the volume: compulsory traffic is tiny and the "
                             "binding roof is the FP32 pipe, see alu"},
        "alu": {"achieved": ALU_SLOTS * B / (k_ms / 1e3) / 1e12, "unit": "T issue-slots/s (FSUB+FFMA per "
                "channel per window entry: the algorithmic count of SURVEY 8d)", "peak": alu_peak / 1e12,
                "frac": ALU_SLOTS * B / (k_ms / 1e3) / alu_peak,
                "frac_at_observed_clock": ALU_SLOTS * B / (k_ms / 1e3) / (148 * 128 * sm_mhz * 1e6),
                "ssd_form": os.environ.get("DM_SSD_FORM", "auto (dot: |a|^2+|b|^2-2ab, one FFMA per term; "
                                           "the kernel executes half the algorithmic slots)"),
                "kernel_ms_covers": "norm pre-pass + sweep (both twin launches)"},
        "e2e": {"value": e2e_val, "unit": "frame-pairs/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "host": {"numa_node_bound": numa_node, "cpus": len(os.sched_getaffinity(0))},
    }

    if not args.no_volume:
        # volume-output mode (the nn.SpatialMatching module contract): HBM-store bound
        Bv = 2
        v1, v2 = in1[:Bv], f2[:Bv]
        dm.match_volume(v1, v2, MAXH, MAXW, ctx=ctx)
        vs = []
        for _ in range(3):
            dm.match_volume(v1, v2, MAXH, MAXW, ctx=ctx)
            vs.append(ctx.last_kernel_ms())
        torch.cuda.synchronize()
        vms = float(np.mean(vs))
        va = BYTES_VOLUME * Bv / (vms / 1e3) / 1e9
        line["volume_mode"] = {"value": Bv / (vms / 1e3), "unit": "frame-pairs/s", "kernel_ms": vms,
                               "roofline": {"bound": "hbm", "achieved": va, "peak": hbm_gbs, "unit": "GB/s",
                                            "frac": va / hbm_gbs, "kernel": "match_volume_kernel<10>"}}
        # the same with the Minus + SoftMax stages (what the reference's model:forward returns):
        # statistics sweep + volume sweep, timed with CUDA events around the call
        import ctypes
        from depthmatch import _lib as dml
        pv = dm.match_volume(v1, v2, MAXH, MAXW, softmax=True, ctx=ctx)   # also the output buffer below
        pr_ = dml.dm_pair()
        pr_.in1, pr_.in2 = v1.data_ptr(), v2.data_ptr()
        pr_.n_pairs, pr_.channels, pr_.h1, pr_.w1, pr_.h2, pr_.w2 = Bv, C, H1, W1, H, W
        pr_.in1_stride_n, pr_.in1_stride_c, pr_.in1_stride_y = v1.stride(0), v1.stride(1), v1.stride(2)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(3):
            dm.api.check(ctx._lib.dm_match_volume(ctx.handle, ctypes.byref(pr_), MAXH, MAXW, dml.DM_VOLUME_NEG_SOFTMAX,
                                                  ctypes.c_void_p(pv.data_ptr())))
        ev[1].record()
        torch.cuda.synchronize()
        sms = ev[0].elapsed_time(ev[1]) / 3
        del pv
        line["volume_softmax_mode"] = {"value": Bv / (sms / 1e3), "unit": "frame-pairs/s", "ms": sms,
                                       "hbm_frac": BYTES_VOLUME * Bv / (sms / 1e3) / 1e9 / hbm_gbs}

    if not args.no_volume:
        # flow only (index + canvas): no probability is asked for, so the kernel skips the soft-max
        out_f = {"index": out_d["index"], "flow_full": out_d["flow_full"]}
        fs = []
        for _ in range(5):
            dm.match_extract(in1, f2, MAXH, MAXW, canvas=(H, W), want=("index",), ctx=ctx, out=out_f)
            fs.append(ctx.last_kernel_ms())
        torch.cuda.synchronize()
        fms = float(np.mean(fs[1:]))
        line["flow_only_mode"] = {"value": B / (fms / 1e3), "unit": "frame-pairs/s (kernels only)", "kernel_ms": fms,
                                  "outputs": ["index", "flow_full"],
                                  "alu_frac": ALU_SLOTS * B / (fms / 1e3) / alu_peak}

    if not args.no_volume:
        # the same path as a frame *stream* (depth_estimation_api.lua keeps the previous frame's
        # features): every frame crosses PCIe once, pair i = (frame i-1, frame i)
        # synthetic stream: every frame is a window of one textured canvas moving by a few pixels
        # per frame (consecutive frames match inside the 33x33 window, like the pairs above)
        srng = np.random.default_rng(777)
        canvas_t = srng.standard_normal((C, H + 64, W + 64)).astype(np.float32)
        offs = np.clip(np.cumsum(srng.integers(-5, 6, (B + 1, 2)), 0), -28, 28) + 32
        stream_frames = torch.empty((B + 1, C, H, W)).pin_memory()
        for i, (sy, sx) in enumerate(offs):
            stream_frames[i] = torch.from_numpy(canvas_t[:, sy:sy + H, sx:sx + W]
                                                + 0.05 * srng.standard_normal((C, H, W)).astype(np.float32))
        fs = dm.FeatureStream(MAXH, MAXW, C, H, W, batch=B, want=want, device=torch.cuda.current_device())
        fs.prime(stream_frames[0])
        batch_frames = stream_frames[1:]
        for _ in range(2):
            h = fs.push(batch_frames)
        fs.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            h = fs.push(batch_frames)
        h.wait()
        fs.synchronize()
        st = time.perf_counter() - t0
        line["e2e_stream"] = {"value": B * args.e2e_steps / st, "unit": "frame-pairs/s",
                              "h2d_bytes_per_step": 4 * B * C * H * W, "d2h_bytes_per_step": int(d2h),
                              "note": "FeatureStream: frames uploaded once, previous frame resident; rank 0 only"}

    if not args.no_cpu:
        import oracle_lib as O
        O.build()
        cores = os.cpu_count() or 1
        rows = H1   # one whole frame pair (1.7 GB of volume + probabilities on the host), best of 4
        cb = cpu_reference_sample(rows, cores, repeats=4)
        cb2 = cpu_reference_sample(41, 2, repeats=1)
        line["cpu_baseline"] = {"value": 1.0 / cb["s_per_pair"], "unit": "frame-pairs/s", "cores": cores,
                                "kind": "port",
                                "sample": "one whole frame pair (%d output rows), best of 4 passes after one warm-up" % rows,
                                "stages_s_per_pair": {k: cb[k] for k in ("match_s", "softmax_s", "extract_s")},
                                "two_threads_value": 1.0 / cb2["s_per_pair"]}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
