"""Multi-GPU sharding of one event: one process per GPU, geometry and tables replicated, gensteps
partitioned, hits gathered at the end.

The reference has no multi-GPU path; what it has is sequential slicing of an event into launches
with an absolute photon_slot_offset so results equal a single launch (sysrap/SGenstep.h:249-323,
qudarap/QSim.cc:479-528, CSGOptiX/CSGOptiX7.cu:415-419).  Sharding across ranks is the same
mechanism run concurrently: rank r simulates a contiguous genstep range with the absolute photon
offset of that range, so the concatenation of the ranks' hits in rank order IS the single-GPU hit
array (ascending photon index).  The only exchange is the final gather: hit counts, then the hit
records, with torch.distributed (NCCL over NVLink on GPUs, gloo on CPU tensors in the tests).
"""
import numpy as np

from .gensteps import partition_gensteps


def shard_event(gensteps, rank, world_size, input_photons=None):
    """-> (gensteps_r, input_photons_r, photon_offset_r, photon_count_r) for this rank.

    Genstep events are split at genstep granularity, balanced by photon count.  An input-photon
    event (one INPUT_PHOTON genstep) is split by photon range instead - something the reference's
    slicing cannot do - each rank getting its own INPUT_PHOTON genstep."""
    gs = np.ascontiguousarray(gensteps, dtype=np.float32).reshape(-1, 6, 4)
    if input_photons is not None:
        n = len(input_photons)
        lo = n * rank // world_size
        hi = n * (rank + 1) // world_size
        g = gs[:1].copy()
        g.view(np.uint32)[0, 0, 3] = hi - lo
        return g, np.ascontiguousarray(input_photons[lo:hi]), lo, hi - lo
    parts = partition_gensteps(gs, world_size)
    s0, s1, off, cnt = parts[rank]
    return np.ascontiguousarray(gs[s0:s1]), None, off, cnt


def gather_hits(hits, group=None, device=None):
    """all-gather variable-length hit arrays (n_r,4,4) float32 -> the whole event's hits in rank
    order.  `hits` may be a numpy array (gloo / CPU) or a torch tensor on this rank's device (NCCL).
    Two collectives: counts, then the records padded to the maximum count."""
    import torch
    import torch.distributed as dist

    as_numpy = isinstance(hits, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(hits, dtype=np.float32)) if as_numpy else hits
    if device is not None:
        t = t.to(device)
    t = t.reshape(-1, 16).contiguous()
    world = dist.get_world_size(group)
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(max(counts), 1)
    pad = torch.zeros((nmax, 16), dtype=torch.float32, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0).reshape(-1, 4, 4)
    return (out.cpu().numpy() if as_numpy else out), counts
