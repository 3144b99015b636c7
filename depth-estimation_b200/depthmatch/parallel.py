"""Multi-GPU partitioning of the matching path.

The path has no exchange step: frame pairs are independent, and output row y of one
pair only reads frame-2 rows y .. y+maxh-1.  So there are two partitions and one optional
collective:

  * shard_pairs      -- pairs (or camera streams) round-robin over ranks; no collective.
  * row_bands        -- one large pair cut into bands of output rows; each rank reads its
                        band of frame 1 plus the band of frame 2 extended by a (maxh-1)-row
                        halo (re-read from the source, never exchanged).
  * gather_bands     -- the only collective: all_gather of the per-band outputs
                        (NCCL over NVLink on GPUs, gloo in the CPU tests).

One process per GPU; `torch.distributed` is plumbing only.
"""
import os

import numpy as np


def shard_pairs(n_pairs, world_size, rank):
    """Indices of the pairs rank `rank` owns (contiguous blocks, sizes differ by at most 1)."""
    base, extra = divmod(n_pairs, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def row_bands(h1, world_size, maxh, align=1):
    """[(y0, y1, halo_y1)] per rank: output rows [y0, y1) and frame-2 rows [y0, halo_y1).
    `align` keeps band edges on multiples of the largest pyramid ratio for multiscale."""
    units = (h1 + align - 1) // align
    base, extra = divmod(units, world_size)
    bands, y = [], 0
    for r in range(world_size):
        n = (base + (1 if r < extra else 0)) * align
        y1 = min(h1, y + n)
        bands.append((y, y1, y1 + maxh - 1 if y1 > y else y))
        y = y1
    return bands


def band_inputs(in1, in2, band):
    """Views of (in1 [C,H1,W1], in2 [C,H2,W2]) a rank needs for its band: no copies."""
    y0, y1, hy1 = band
    return in1[..., y0:y1, :], in2[..., y0:hy1, :]


def equal_row_bands(h1, world_size, maxh):
    """Bands of EQUAL height hb = ceil(h1 / world) (the last ones may be cut or empty), so that the
    per-band outputs are the rank-th hb-row slices of one [world*hb, ...] buffer and one in-place
    all_gather_into_tensor assembles the full map with no staging copy.  Returns (hb, bands)."""
    hb = (h1 + world_size - 1) // world_size
    bands = []
    for r in range(world_size):
        y0, y1 = min(h1, r * hb), min(h1, (r + 1) * hb)
        bands.append((y0, y1, y1 + maxh - 1 if y1 > y0 else y0))
    return hb, bands


def gather_bands_inplace(full, hb, rank, dist=None, async_op=False):
    """The only collective of the path.  `full` is the preallocated [world*hb, ...] map (rows along
    dim 0) whose slice [rank*hb, (rank+1)*hb) this rank has just written: one
    all_gather_into_tensor, input aliasing its own slot of the output (NCCL's in-place form)."""
    if dist is None:
        import torch.distributed as dist
    mine = full.narrow(0, rank * hb, hb)
    return dist.all_gather_into_tensor(full, mine, async_op=async_op)


def gather_bands(local, bands, dist=None, dim=0):
    """all_gather per-band outputs (torch tensors, band rows along `dim`) of UNEQUAL bands
    (row_bands with an alignment) into the full map on every rank: pad to the tallest band."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    hmax = max(b[1] - b[0] for b in bands)
    pad_shape = list(local.shape)
    pad_shape[dim] = hmax
    padded = local.new_zeros(pad_shape)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p.narrow(dim, 0, b[1] - b[0]) for p, b in zip(parts, bands)], dim=dim)


class RowBandMatcher:
    """BASELINE config 5: a stream of single large frame pairs, each cut in `world` equal row bands
    (SURVEY 8e).  Per pair a rank runs the fused kernel on its band -- frame-1 rows [y0, y1),
    frame-2 rows [y0, y1 + maxh - 1): the halo is re-read from the source, never exchanged -- with
    the kernel writing straight into this rank's slice of the final maps; then the band is handed
    to the other ranks on a side stream, under the next pair's sweep (two result sets alternate).

    Two ways to hand the band over (`gather`):
      "peer"  the final maps live in symmetric memory (torch.distributed._symmetric_memory: every
              rank maps every other rank's buffer over NVLink / NVSwitch); a rank copies its band
              slice into each peer's map with device-to-device copies, which the copy engines
              execute -- the sweep is a persistent kernel that owns every SM's register file, an
              SM-resident collective kernel can only run in the gaps between sweeps -- followed by
              one symmetric-memory barrier.
      "nccl"  one in-place all_gather_into_tensor per output (gather_bands_inplace).
    "auto" tries "peer" and falls back to "nccl" (no P2P mapping, single rank, CPU tests).

        m = RowBandMatcher(dm, h1, w1, maxh, maxw, rank, world, dist, want=("index", "pmax"))
        h = m.step(in1, in2)      # torch CUDA tensors [C,H1,W1], [C,H2,W2], every rank holds them
        maps = h.wait()           # {"index": [h1, w1], ...} on every rank
    """

    def __init__(self, dm, h1, w1, maxh, maxw, rank, world, dist=None, want=("index", "pmax"), ctx=None,
                 gather="auto"):
        import torch
        self.dm, self.maxh, self.maxw, self.rank, self.world, self.dist = dm, maxh, maxw, rank, world, dist
        self.h1, self.w1, self.want = h1, w1, tuple(want)
        self.hb, self.bands = equal_row_bands(h1, world, maxh)
        dt = {"index": torch.int64, "index_thr": torch.int64}
        shape = (world * self.hb, w1)
        self.gather, self.sym, self.peers = "none", None, None
        if world > 1 and gather in ("auto", "peer"):
            try:
                self._setup_peer(torch, shape, dt)
                self.gather = "peer"
            except Exception as e:  # no symmetric memory on this box / build: NCCL does the gather
                if gather == "peer":
                    raise
                self.gather_fallback_reason = "%s: %s" % (type(e).__name__, str(e)[:200])
        if self.gather != "peer":
            self.full = [{k: torch.zeros(shape, dtype=dt.get(k, torch.float32), device="cuda") for k in self.want}
                         for _ in range(2)]
            if world > 1:
                self.gather = "nccl"
        self.ctx = ctx
        self.side = torch.cuda.Stream() if world > 1 else None
        self.gathered = [None, None]   # event: the hand-over that last wrote result set i is complete
        self.i = 0

    def _setup_peer(self, torch, shape, dt):
        import torch.distributed as tdist
        import torch.distributed._symmetric_memory as symm_mem
        group = tdist.group.WORLD
        self.full, self.sym, self.peers = [], [], []
        for _ in range(2):
            maps, hdls, views = {}, {}, {}
            for k in self.want:
                t = symm_mem.empty(shape, dtype=dt.get(k, torch.float32), device=torch.device("cuda", torch.cuda.current_device()))
                t.zero_()
                hdl = symm_mem.rendezvous(t, group)
                maps[k], hdls[k] = t, hdl
                views[k] = [hdl.get_buffer(r, shape, t.dtype) if r != self.rank else t for r in range(self.world)]
            self.full.append(maps)
            self.sym.append(hdls)
            self.peers.append(views)
        torch.cuda.synchronize()

    class _Handle:
        def __init__(self, maps, event, h1):
            self.maps, self.event, self.h1 = maps, event, h1

        def wait(self):
            """Orders the caller's current stream after the hand-over and returns the full maps."""
            import torch
            if self.event is not None:
                torch.cuda.current_stream().wait_event(self.event)
            return {k: v[:self.h1] for k, v in self.maps.items()}

    def step(self, in1, in2):
        import torch
        s = self.i & 1
        self.i += 1
        full = self.full[s]
        cur = torch.cuda.current_stream()
        if self.gathered[s] is not None:
            cur.wait_event(self.gathered[s])   # the hand-over two steps ago still reads / writes this set
        y0, y1, hy1 = self.bands[self.rank]
        if y1 > y0:
            a, b = band_inputs(in1, in2, self.bands[self.rank])
            out = {k: full[k][y0:y1].unsqueeze(0) for k in self.want}
            self.dm.match_extract(a, b, self.maxh, self.maxw, want=self.want, out=out, ctx=self.ctx)
        if self.world == 1:
            return self._Handle(full, None, self.h1)
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(done)
            if self.gather == "peer":
                if y1 > y0:
                    for k in self.want:
                        src = full[k][y0:y1]
                        for r in range(self.world):
                            if r != self.rank:
                                self.peers[s][k][r][y0:y1].copy_(src, non_blocking=True)
                # every rank's copies are stream-ordered before its arrival at the barrier: past it,
                # all bands have landed in this rank's maps
                self.sym[s][self.want[0]].barrier(channel=s)
            else:
                for k in self.want:
                    gather_bands_inplace(full[k], self.hb, self.rank, self.dist)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.gathered[s] = ev
        return self._Handle(full, ev, self.h1)


def match_extract_row_bands(dm, in1, in2, maxh, maxw, rank, world_size, dist=None, **kw):
    """One large pair, one call (no pipelining): this rank's band through the fused kernel, then the
    in-place gather.  in1/in2 are torch CUDA tensors [C,H,W]."""
    want = tuple(k for k in kw.pop("want", ("index", "pmax")) if k not in ("n_untouched", "flow_full"))
    m = RowBandMatcher(dm, in1.shape[-2], in1.shape[-1], maxh, maxw, rank, world_size, dist, want=want, **kw)
    return m.step(in1, in2).wait()


def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device=0):
    """One process per GPU: run this process (and so the first-touch placement of the pinned
    staging buffers it allocates afterwards) on the CPUs of the NUMA node the GPU's PCIe root
    hangs off.  With several ranks streaming host buffers at PCIe rate, remote-node buffers
    halve the aggregate.  Returns the node, or None when the topology is not exposed."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, AttributeError, ValueError, RuntimeError):
        return None


def numa_report(device=0):
    """What the box exposes about the GPU's NUMA placement (why bind_to_gpu_numa_node may return
    None): the sysfs numa_node of the GPU's PCI function (-1 = the hypervisor hides it) and the
    number of NUMA nodes the OS sees."""
    rep = {"gpu_numa_node_sysfs": None, "nodes_online": None}
    try:
        import torch
        prop = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            rep["gpu_numa_node_sysfs"] = int(f.read().strip())
    except (OSError, AttributeError, ValueError, RuntimeError):
        pass
    try:
        with open("/sys/devices/system/node/online") as f:
            rep["nodes_online"] = f.read().strip()
    except OSError:
        pass
    return rep
