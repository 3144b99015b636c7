"""BASELINE config 5: a 1920x1080 frame-pair stream, 65x65 window, row bands with a 64-row halo
across the GPUs of one box, one NCCL all_gather of the band outputs per pair.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29540 scripts/bench_c5_bands.py [--pairs 8]

Every rank holds the pair's feature maps (synthetic, same seed), computes its band through the
fused kernel and gathers index + pmax; time = CUDA events around the whole loop, max over
ranks.  Rank 0 prints one JSON line; it also checks the gathered map against the planted flow."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
import depthmatch as dm  # noqa: E402
from depthmatch import parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    C, H, W, mh = 10, 1080, 1920, 65
    g = torch.Generator(device="cuda").manual_seed(7)
    in2 = torch.randn((C, H, W), device="cuda", generator=g)
    fy, fx = 9, -14                                   # planted constant flow
    c = mh // 2
    in1 = in2[:, c + fy:c + fy + H - mh + 1, c + fx:c + fx + W - mh + 1].contiguous()
    in1 += 0.05 * torch.randn(in1.shape, device="cuda", generator=g)
    want = ("index", "pmax")

    def step():
        if world > 1:
            return parallel.match_extract_row_bands(dm, in1, in2, mh, mh, rank, world, dist, want=want)
        return dm.match_extract(in1, in2, mh, mh, want=want)

    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.pairs):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        idx = out["index"]
        ok = bool(((idx - 1) // mh == c + fy).all() and ((idx - 1) % mh == c + fx).all())
        slots = 2.0 * C * mh * mh * (H - mh + 1) * (W - mh + 1)
        per_pair = float(ms.item()) / args.pairs
        print(json.dumps({"config": "c5: 1920x1080, 65x65, C=10, %d row bands + all_gather" % world,
                          "n_gpus": world, "pairs": args.pairs, "ms_per_pair": per_pair,
                          "pairs_per_s": 1e3 / per_pair,
                          "alu_frac_aggregate": slots / (per_pair * 1e-3) / (world * 148 * 128 * 1.965e9),
                          "gathered_shape": list(idx.shape), "planted_flow_recovered": ok}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
