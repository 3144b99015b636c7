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


def gather_bands(local, bands, dist=None, dim=0):
    """all_gather per-band outputs (torch tensors, band rows along `dim`) into the full
    map on every rank.  Bands may differ in height by one unit: pad to the tallest."""
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


def match_extract_row_bands(dm, in1, in2, maxh, maxw, rank, world_size, dist=None, **kw):
    """Config-5 style execution of one large pair: this rank's band through the fused kernel,
    then one all_gather of index / score maps.  in1/in2 are torch CUDA tensors [C,H,W]."""
    bands = row_bands(in1.shape[-2], world_size, maxh)
    a, b = band_inputs(in1, in2, bands[rank])
    out = dm.match_extract(a, b, maxh, maxw, **kw)
    return {k: gather_bands(v, bands, dist, dim=v.dim() - 2) for k, v in out.items()
            if k not in ("n_untouched", "flow_full")}


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
