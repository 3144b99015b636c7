"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes check the pair
sharding, the row-band + halo partition and the all_gather of band outputs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
    from depthmatch import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H1, W1, maxh = 37, 11, 9
        in1 = torch.arange(2 * H1 * W1, dtype=torch.float32).reshape(2, H1, W1)
        in2 = torch.arange(2 * (H1 + maxh - 1) * W1, dtype=torch.float32).reshape(2, H1 + maxh - 1, W1)
        bands = parallel.row_bands(H1, world, maxh)
        a, b = parallel.band_inputs(in1, in2, bands[rank])
        # stand-in for the kernel: something that needs exactly the halo rows
        local = torch.stack([b[0, y:y + maxh].sum(0) + a[1, y] for y in range(a.shape[1])])
        full = parallel.gather_bands(local, bands, dist, dim=0)
        want = torch.stack([in2[0, y:y + maxh].sum(0) + in1[1, y] for y in range(H1)])
        ok = bool(torch.equal(full, want))
        # config-5 form: equal bands, the "kernel" writes its slice of the final map, in-place gather
        hb, eb = parallel.equal_row_bands(H1, world, maxh)
        final = torch.zeros((world * hb, W1))
        y0, y1, hy1 = eb[rank]
        a, b = parallel.band_inputs(in1, in2, eb[rank])
        assert b.shape[1] == y1 - y0 + maxh - 1
        final[y0:y1] = torch.stack([b[0, y:y + maxh].sum(0) + a[1, y] for y in range(y1 - y0)])
        parallel.gather_bands_inplace(final, hb, rank, dist)
        ok = ok and bool(torch.equal(final[:H1], want))
        mine = list(parallel.shard_pairs(7, world, rank))
        got = [None] * world
        dist.all_gather_object(got, mine)
        ok = ok and sorted(sum(got, [])) == list(range(7))
        t = torch.tensor([1.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # bench.py's max-over-ranks timing
        ok = ok and t.item() == float(world)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_band_partition_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("h1,world,maxh,align", [(328, 8, 33, 1), (1016, 8, 65, 1), (5, 8, 3, 1),
                                                 (360, 3, 8, 4)])
def test_row_bands_cover_exactly(h1, world, maxh, align):
    sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
    from depthmatch import parallel
    bands = parallel.row_bands(h1, world, maxh, align)
    assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == h1
    for (a0, a1, ah), (b0, b1, bh) in zip(bands, bands[1:]):
        assert a1 == b0
    for y0, y1, hy in bands:
        assert hy == (y1 + maxh - 1 if y1 > y0 else y0)
        assert y0 % align == 0


@pytest.mark.parametrize("h1,world,maxh", [(1016, 8, 65), (328, 8, 33), (37, 2, 9), (5, 8, 3), (1016, 1, 65)])
def test_equal_row_bands(h1, world, maxh):
    sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
    from depthmatch import parallel
    hb, bands = parallel.equal_row_bands(h1, world, maxh)
    assert hb * world >= h1 and (hb - 1) * world < h1 and len(bands) == world
    rows = [y for y0, y1, _ in bands for y in range(y0, y1)]
    assert rows == list(range(h1))
    for r, (y0, y1, hy) in enumerate(bands):
        assert y0 == min(h1, r * hb) and y1 - y0 <= hb and hy == (y1 + maxh - 1 if y1 > y0 else y0)


def test_shard_pairs_balanced():
    sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
    from depthmatch import parallel
    for n, w in ((64, 8), (7, 2), (3, 8), (0, 4)):
        parts = [list(parallel.shard_pairs(n, w, r)) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
