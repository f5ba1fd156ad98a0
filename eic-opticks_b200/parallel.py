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


def gather_hits(hits, group=None, device=None, dst=None):
    """gather variable-length hit arrays (n_r,4,4) float32 -> the whole event's hits in rank order.
    `hits` may be a numpy array (gloo / CPU) or a torch tensor on this rank's device (NCCL).

    dst=None : every rank receives the whole array (all-gather of the records padded to the maximum count).
    dst=r    : only rank r receives it (the reference hands hits to ONE host process): every other rank sends its
               records straight into its slice of r's output buffer (point-to-point, no padding, no extra copy) and
               gets None back.
    Returns (hits or None, counts per rank)."""
    import torch
    import torch.distributed as dist

    as_numpy = isinstance(hits, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(hits, dtype=np.float32)) if as_numpy else hits
    if device is not None:
        t = t.to(device)
    t = t.reshape(-1, 16).contiguous()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    all_counts = torch.zeros(world, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(all_counts, n_local, group=group)
    counts = [int(c) for c in all_counts.tolist()]
    if dst is not None:
        out = None
        ops = []
        if rank == dst:
            out = torch.empty((sum(counts), 16), dtype=torch.float32, device=t.device)
            off = 0
            for r, c in enumerate(counts):
                if r == rank:
                    out[off:off + c] = t
                elif c:
                    ops.append(dist.P2POp(dist.irecv, out[off:off + c], dist.get_global_rank(group, r) if group is not None else r, group))
                off += c
        elif t.shape[0]:
            ops.append(dist.P2POp(dist.isend, t, dist.get_global_rank(group, dst) if group is not None else dst, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if out is None:
            return None, counts
        out = out.reshape(-1, 4, 4)
        return (out.cpu().numpy() if as_numpy else out), counts
    nmax = max(max(counts), 1)
    pad = torch.zeros((nmax, 16), dtype=torch.float32, device=t.device)
    pad[: t.shape[0]] = t
    buf = torch.empty((world * nmax, 16), dtype=torch.float32, device=t.device)      # concatenation layout (gloo and NCCL both take it)
    dist.all_gather_into_tensor(buf, pad, group=group)
    out = torch.cat([buf[r * nmax: r * nmax + c] for r, c in enumerate(counts)], dim=0).reshape(-1, 4, 4)
    return (out.cpu().numpy() if as_numpy else out), counts
