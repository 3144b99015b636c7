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
on this box's host cores on a bounded sample; `parity` compares pair 0 of the timed batch with the
oracle (outside the timed region); `hard_data` repeats the measurement on unrelated frames;
`row_bands` is BASELINE config 5 (one 1920x1080 / 65x65 pair cut in N row bands + the in-place
NCCL gather, strong scaling) at the same N.
"""
import argparse
import contextlib
import io
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
# dram__bytes_read.sum + dram__bytes_write.sum of the sweep kernel per launch / pairs per launch, from
# this round's `ncu --set full` capture (file named next to it; None until a capture of this build exists)
NCU_TRAFFIC = {"bytes_per_pair": (73.1e6 + 5.1e6) / 4, "source": "profiles/r02_ncu_kernels.md, launch #2 (4 pairs)"}
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
    flows = []
    for b in range(B):
        in1, in2, flow = make_pair(C, H, W, MAXH, MAXW, seed=seed0 + b, noise=0.05)
        f1[b] = np.random.default_rng(seed0 + b + 7919).standard_normal((C, H, W), dtype=np.float32)
        f1[b, :, cy:cy + H1, cx:cx + W1] = in1
        f2[b] = in2
        flows.append(flow)
    return f1, f2, flows


def crop(f1):
    cy, cx = (MAXH + 1) // 2 - 1, (MAXW + 1) // 2 - 1
    return f1[:, :, cy:cy + H1, cx:cx + W1]


def cpu_reference_sample(rows, nthreads, repeats=2, warm=True):
    """The reference CPU path (oracle port: SpatialMatching -> Minus -> SoftMax -> argmax+tie ->
    extractOutput(0.11) -> canvas) on `rows` output rows of one north pair.  Returns s/pair."""
    import oracle_lib as O
    from synth import make_pair
    in1, in2, _ = make_pair(C, H, W, MAXH, MAXW, seed=1234, noise=0.05)
    a, b = in1[:, :rows], in2[:, :rows + MAXH - 1]
    best = None
    for it in range(0 if warm else 1, repeats + 1):
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


WANT = ("index", "pmax", "score_thr")


def run_config(B):
    """`config` of both arms (the reference arm times whole pairs of the same workload one by one)."""
    return {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "mode": "fused",
            "outputs": list(WANT) + ["flow_full"],
            "l2": "inputs per step (%.0f MB) larger than the 126 MB L2" % (B * 2 * C * H * W * 4 / 1e6)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference itself needs Torch7/Lua which this image does not have), all host threads."""
    if rank != 0:
        return
    import oracle_lib as O
    O.build()
    cores = os.cpu_count() or 1
    rows = H1  # one WHOLE frame pair per step (0.6 s on 32 cores, ~3 s on 8)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_sample(rows, cores, repeats=1, warm=False)
    t0 = time.perf_counter()
    per = [cpu_reference_sample(rows, cores, repeats=1, warm=False)["s_per_pair"] for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    s_pair = float(np.mean(per))
    val = 1.0 / s_pair
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frame-pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": run_config(args.batch),
            "cpu_baseline": {"value": val, "unit": "frame-pairs/s", "cores": cores, "kind": "port",
                             "sample": "one whole frame pair (all %d output rows) per step, all host threads; "
                                       "the reference itself needs Torch7/Lua, so this is the oracle port" % rows},
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
    ap.add_argument("--no-volume", action="store_true", help="skip the side measurements (volume, flow-only, "
                    "stream, hard data)")
    ap.add_argument("--no-bands", action="store_true", help="skip the config-5 row-band record")
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
    f1_h, f2_h, flows = make_batch(B, 1234 + rank * 1000)
    f1 = torch.from_numpy(f1_h).cuda()
    f2 = torch.from_numpy(f2_h).cuda()
    in1 = crop(f1)
    want = WANT

    def gather_list(x):
        """per-rank scalars -> list on every rank"""
        t = torch.tensor([float(x)], device="cuda")
        if dist is None:
            return [float(x)]
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [float(v.item()) for v in parts]

    def max_over_ranks(x):
        return max(gather_list(x))

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
    ms_total = max_over_ranks(ms_total)
    value = world * B * args.steps / (ms_total / 1e3)
    rescored, exact_pass = ctx.last_counts()

    # the planted flow came back, on every pair of this rank's batch (outside the timed region)
    idx_all = out["index"].cpu().numpy()
    cyc = (MAXH + 1) // 2
    for b in range(B):
        assert np.array_equal((idx_all[b] - 1) // MAXW + 1 - cyc, flows[b][0]), "planted y-flow not recovered, pair %d" % b
        assert np.array_equal((idx_all[b] - 1) % MAXW + 1 - cyc, flows[b][1]), "planted x-flow not recovered, pair %d" % b
    got0 = {k: v[0].cpu().numpy() for k, v in out.items()}   # pair 0 for the parity block

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
        e2e_step(i)
    for c in ctxs:
        c.synchronize()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # the host-buffer path returned what the device-buffer path computed (outside the timed region)
    for o in outs:
        assert np.array_equal(o["index"], idx_all), "e2e result differs from the device path"
    e2e_ranks = gather_list(e2e_s)
    e2e_val = world * B * args.e2e_steps / max(e2e_ranks)
    h2d = 4 * (B * C * H1 * W1 + B * C * H * W)   # the frame-1 crop is packed by a 3-D copy
    d2h = B * (H1 * W1 * (8 + 4 + 4) + 2 * H * W * 4)

    # ---- the host-copy ceiling of this box: every rank copies pinned buffers of the e2e step's size
    # in both directions AT THE SAME TIME (two streams), nothing else running
    cp_in = torch.empty(h2d // 4, dtype=torch.float32).pin_memory()
    cp_out = torch.empty(d2h // 4, dtype=torch.float32).pin_memory()
    d_in, d_out = torch.empty_like(cp_in, device="cuda"), torch.empty_like(cp_out, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def copy_round():
        with torch.cuda.stream(s_in):
            d_in.copy_(cp_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            cp_out.copy_(d_out, non_blocking=True)

    copy_round()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        copy_round()
    torch.cuda.synchronize()
    cp_s = (time.perf_counter() - t0) / 4
    cp_ranks = gather_list(cp_s)
    del cp_in, cp_out, d_in, d_out

    # ---- the same path as a frame *stream* (depth_estimation_api.lua keeps the previous frame's
    # features): every frame crosses PCIe once, pair i = (frame i-1, frame i); every rank runs its own
    # stream.  Synthetic: every frame is a window of one textured canvas moving by a few pixels per
    # frame (consecutive frames match inside the 33x33 window, like the pairs above)
    stream_val = None
    if not args.no_volume:
        srng = np.random.default_rng(777 + rank)
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
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            h = fs.push(batch_frames)
        h.wait()
        fs.synchronize()
        st_ranks = gather_list(time.perf_counter() - t0)
        stream_val = {"value": world * B * args.e2e_steps / max(st_ranks), "unit": "frame-pairs/s",
                      "h2d_bytes_per_step": 4 * B * C * H * W, "d2h_bytes_per_step": int(d2h),
                      "seconds_per_rank": st_ranks,
                      "note": "FeatureStream on every rank: frames uploaded once, previous frame resident"}
        del fs, stream_frames

    # ---- BASELINE config 5 at this N: one 1920x1080 pair, 65x65 window, N equal row bands with a
    # 64-row halo, in-place NCCL all_gather of index + pmax on a side stream (strong scaling)
    bands_rec = None
    if not args.no_bands:
        bands_rec = bench_row_bands(dm, dm_parallel, dist, rank, world, ctx)

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
    exec_slots = (C + 1) * K * H1 * W1   # dot form: C FFMA + the norm-sum FADD per window entry
    line = {
        "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": run_config(B),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": achieved / hbm_gbs,
                     "traffic": NCU_TRAFFIC["bytes_per_pair"] * B, "traffic_unit": "bytes per launch",
                     "traffic_source": NCU_TRAFFIC["source"],
                     "algorithmic_bytes": BYTES_FUSED * B, "peak_source": which,
                     "kernel": "match_extract_kernel<10>", "kernel_ms": k_ms,
                     "note": "fused mode never writes the volume: compulsory traffic is tiny and the "
                             "binding roof is the FP32 pipe, see alu"},
        "alu": {"achieved": ALU_SLOTS * B / (k_ms / 1e3) / 1e12, "unit": "T issue-slots/s (FSUB+FFMA per "
                "channel per window entry: the algorithmic count of SURVEY 8d)", "peak": alu_peak / 1e12,
                "frac": ALU_SLOTS * B / (k_ms / 1e3) / alu_peak,
                "frac_at_observed_clock": ALU_SLOTS * B / (k_ms / 1e3) / (148 * 128 * sm_mhz * 1e6),
                # what the kernel actually executes: the dot form needs (C + 1) lane-operations per
                # window entry instead of 2C, so this is the fraction of the FP32 pipe really in use
                "frac_executed": exec_slots * B / (k_ms / 1e3) / alu_peak,
                "ssd_form": "auto -> dot (|a|^2+|b|^2-2ab, one FFMA per term; pixels it cannot order are "
                            "rescored entry by entry)",
                "kernel_ms_covers": "norm pre-pass + sweep (both twin launches)"},
        "e2e": {"value": e2e_val, "unit": "frame-pairs/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h),
                "per_rank": {"seconds": e2e_ranks,
                             "h2d_gbs": [h2d * args.e2e_steps / t / 1e9 for t in e2e_ranks],
                             "d2h_gbs": [d2h * args.e2e_steps / t / 1e9 for t in e2e_ranks]},
                # every rank copying the same bytes both ways at once, no kernels: what the box's host
                # memory / PCIe path gives N ranks together -- the ceiling of any e2e number at this N
                "host_copy_ceiling": {"pairs_per_s": world * B / max(cp_ranks),
                                      "h2d_gbs_per_rank": [h2d / t / 1e9 for t in cp_ranks],
                                      "d2h_gbs_per_rank": [d2h / t / 1e9 for t in cp_ranks],
                                      "aggregate_gbs": sum((h2d + d2h) / t / 1e9 for t in cp_ranks)}},
        "gpu_launches": int(launches),
        "near_ties_logged": {"rescored_pixels_last_step": rescored, "exact_pass_pixels_last_step": exact_pass,
                             "pixels_per_step": B * H1 * W1},
        "clocks": clocks,
        "host": {"numa_node_bound": numa_node, "numa": dm_parallel.numa_report(local),
                 "cpus": len(os.sched_getaffinity(0))},
    }
    if stream_val:
        line["e2e_stream"] = stream_val
    if bands_rec:
        line["row_bands"] = bands_rec

    if not args.no_cpu:
        # ---- parity of the timed configuration: pair 0 of the batch against the oracle
        import oracle_lib as O
        from parity import oracle_pair, parity_report
        O.build()
        want0 = oracle_pair(O, np.ascontiguousarray(crop(f1_h)[0]), f2_h[0], MAXH, MAXW, canvas=(H, W))
        rep = parity_report(got0, want0)
        rep["what"] = "pair 0 of the timed batch (auto form = dot) vs the CPU oracle; bars: indices bit-exact " \
                      "outside near-ties (top-2 gap < 1e-5), scores within 1e-4"
        line["parity"] = rep
        del want0

    if not args.no_volume:
        # ---- hard data: two unrelated frames per pair (no match anywhere: flat soft-max, most pixels
        # near the 0.11 threshold) -- the data-dependent side of the thresholded output
        hrng = np.random.default_rng(99)
        Bh = min(B, 4)
        h1_t = torch.from_numpy(hrng.standard_normal((Bh, C, H1, W1), dtype=np.float32)).cuda()
        h2_t = torch.from_numpy(hrng.standard_normal((Bh, C, H, W), dtype=np.float32)).cuda()
        out_h = {k: v[:Bh] for k, v in out_d.items()}
        for _ in range(3):
            outh = dm.match_extract(h1_t, h2_t, MAXH, MAXW, canvas=(H, W), want=want, ctx=ctx, out=out_h)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dm.match_extract(h1_t, h2_t, MAXH, MAXW, canvas=(H, W), want=want, ctx=ctx, out=out_h)
        e1.record()
        torch.cuda.synchronize()
        hms = e0.elapsed_time(e1) / 5
        hres, hex_ = ctx.last_counts()
        line["hard_data"] = {"value": Bh / (hms / 1e3), "unit": "frame-pairs/s", "ms_per_step": hms,
                             "pairs_per_step": Bh, "data": "unrelated N(0,1) frames",
                             "rescored_pixels": hres, "exact_pass_pixels": hex_, "pixels": Bh * H1 * W1}
        if not args.no_cpu:
            wanth = oracle_pair(O, h1_t[0].cpu().numpy(), h2_t[0].cpu().numpy(), MAXH, MAXW, canvas=(H, W))
            line["hard_data"]["parity"] = parity_report({k: v[0].cpu().numpy() for k, v in outh.items()}, wanth)
            del wanth
        del h1_t, h2_t

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
                                            "frac": va / hbm_gbs, "kernel": "match_volume_px_kernel<10, kFma>",
                                            "traffic": 1729.1e6, "traffic_unit": "bytes per launch of 2 pairs (1679.6 MB written + 49.5 MB read; "
                                            "the last ~57 MB of the volume are still dirty in L2 when the kernel ends)",
                                            "traffic_source": "profiles/r02_ncu_volume_strip.md, launch #16", "algorithmic_bytes": BYTES_VOLUME * Bv}}
        # the round-1 kernel (sector stores from the tiled sweep), for the record
        ctx.set_option("volume_kernel", 1)
        vs1 = []
        for _ in range(3):
            dm.match_volume(v1, v2, MAXH, MAXW, ctx=ctx)
            vs1.append(ctx.last_kernel_ms())
        torch.cuda.synchronize()
        ctx.set_option("volume_kernel", 0)
        line["volume_mode"]["tiled_kernel_ms"] = float(np.mean(vs1[1:]))
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

        # flow only (index + canvas): no probability is asked for, so the kernel skips the soft-max
        out_f = {"index": out_d["index"], "flow_full": out_d["flow_full"]}
        fs_ = []
        for _ in range(5):
            dm.match_extract(in1, f2, MAXH, MAXW, canvas=(H, W), want=("index",), ctx=ctx, out=out_f)
            fs_.append(ctx.last_kernel_ms())
        torch.cuda.synchronize()
        fms = float(np.mean(fs_[1:]))
        line["flow_only_mode"] = {"value": B / (fms / 1e3), "unit": "frame-pairs/s (kernels only)", "kernel_ms": fms,
                                  "outputs": ["index", "flow_full"],
                                  "alu_frac": ALU_SLOTS * B / (fms / 1e3) / alu_peak,
                                  "alu_frac_executed": exec_slots * B / (fms / 1e3) / alu_peak}

    if not args.no_cpu:
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


def bench_row_bands(dm, dm_parallel, dist, rank, world, ctx, pairs=12, warmup=3):
    """One 1920x1080 / 65x65 / C=10 pair per step over `world` row bands (RowBandMatcher): every rank
    holds the pair, sweeps its band into its slice of the final maps, the in-place all_gather runs on
    a side stream under the next pair's sweep.  CUDA events around `pairs` steps, max over ranks."""
    import torch
    Cb, Hb, Wb, mh = 10, 1080, 1920, 65
    g = torch.Generator(device="cuda").manual_seed(7)
    in2 = torch.randn((Cb, Hb, Wb), device="cuda", generator=g)
    fy, fx = 9, -14                                   # planted constant flow
    c = mh // 2
    h1, w1 = Hb - mh + 1, Wb - mh + 1
    in1 = in2[:, c + fy:c + fy + h1, c + fx:c + fx + w1].contiguous()
    in1 += 0.05 * torch.randn(in1.shape, device="cuda", generator=g)
    m = dm_parallel.RowBandMatcher(dm, h1, w1, mh, mh, rank, world, dist, want=("index", "pmax"), ctx=ctx)
    for _ in range(warmup):
        h = m.step(in1, in2)
    h.wait()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(pairs):
        h = m.step(in1, in2)
    maps = h.wait()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    # the sweep alone and the gather alone, one pair, for the breakdown
    a, b = dm_parallel.band_inputs(in1, in2, m.bands[rank])
    ctx.set_profiling(True)
    dm.match_extract(a, b, mh, mh, want=("index", "pmax"), ctx=ctx)
    sweep_ms = torch.tensor([ctx.last_kernel_ms()], device="cuda")
    gather_ms = torch.zeros(1, device="cuda")
    if dist is not None:
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for k in ("index", "pmax"):   # for the breakdown: what one NCCL gather of the same bytes costs when nothing overlaps
            dm_parallel.gather_bands_inplace(m.full[0][k], m.hb, rank, dist)
        g1.record()
        torch.cuda.synchronize()
        gather_ms[0] = g0.elapsed_time(g1)
        for t in (ms, sweep_ms, gather_ms):
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
    idx = maps["index"]
    ok = bool(((idx - 1) // mh == c + fy).all() and ((idx - 1) % mh == c + fx).all()) and tuple(idx.shape) == (h1, w1)
    per_pair = float(ms.item()) / pairs
    slots = 2.0 * Cb * mh * mh * h1 * w1
    return {"config": "c5: 1920x1080, 65x65, C=10, %d row band(s) of %d rows + 64-row halo, band outputs (index, pmax) "
                      "handed to the other ranks on a side stream" % (world, m.hb),
            "gather": m.gather, "gather_fallback_reason": getattr(m, "gather_fallback_reason", None),
            "scaling": "strong", "n_gpus": world, "pairs": pairs, "ms_per_pair": per_pair,
            "pairs_per_s": 1e3 / per_pair, "band_sweep_ms": float(sweep_ms.item()),
            "nccl_gather_ms_if_serial": float(gather_ms.item()),
            "gather_bytes_per_rank": int(m.hb * w1 * 12),
            "alu_frac_aggregate": slots / (per_pair * 1e-3) / (world * 148 * 128 * 1.965e9),
            "planted_flow_recovered": ok}


if __name__ == "__main__":
    # stdout carries exactly one JSON line.  Libraries write there too (NCCL's version banner comes through
    # C stdio, whatever NCCL_DEBUG_FILE says), so file descriptor 1 points at stderr while the bench runs and
    # the line goes to the real stdout at the end.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(_buf):
            main()
    finally:
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        os.close(_real_stdout)
        sys.stdout.write(_buf.getvalue())
        sys.stdout.flush()
