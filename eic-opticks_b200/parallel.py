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


class PipelinedHitGather:
    """Root-only gather of each event's hits, overlapped with the following events (GPU ranks, NCCL).

    push(sim) after every simulate: the event's hits are copied into one of two staging buffers on the simulator's
    stream, and the gather of the PREVIOUS event is posted on a second stream - by then every rank has long finished that
    event, so the exchange of the counts does not make a fast rank wait for a slow one, and the records travel while the
    next event's kernels run.  drain() posts the last gather and waits for it.  The root's receive buffer only grows.
    result() -> (hits (n,4,4) device tensor on the root | None, counts per rank) of the last completed gather."""

    def __init__(self, capacity, device, dst=0, group=None):
        import torch
        self.torch = torch
        self.device, self.dst, self.group = device, dst, group
        self.stream = torch.cuda.Stream(device)
        self.bufs = [torch.empty((max(int(capacity), 1), 16), dtype=torch.float32, device=device) for _ in range(2)]
        self.staged = None              # (buffer index, hit count, event recorded after the staging copy)
        self.k = 0
        self.out = None
        self.last = (None, None)

    def _post(self):
        import torch.distributed as dist
        torch = self.torch
        b, n, ev = self.staged
        self.staged = None
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            t = self.bufs[b][:n]
            n_local = torch.tensor([n], dtype=torch.int64, device=self.device)
            all_counts = torch.zeros(world, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(all_counts, n_local, group=self.group)
            counts = [int(c) for c in all_counts.tolist()]
            ops, out = [], None
            if rank == self.dst:
                tot = sum(counts)
                if self.out is None or self.out.shape[0] < tot:
                    self.out = torch.empty((tot + tot // 4 + 1024, 16), dtype=torch.float32, device=self.device)
                out = self.out[:tot]
                off = 0
                for r, c in enumerate(counts):
                    if r == rank:
                        out[off:off + c].copy_(t, non_blocking=True)
                    elif c:
                        ops.append(dist.P2POp(dist.irecv, out[off:off + c], dist.get_global_rank(self.group, r) if self.group is not None else r, self.group))
                    off += c
            elif n:
                ops.append(dist.P2POp(dist.isend, t, dist.get_global_rank(self.group, self.dst) if self.group is not None else self.dst, self.group))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()                  # orders the gather stream after the transfer; the host does not block
            self.last = (None if out is None else out.reshape(-1, 4, 4), counts)

    def push(self, sim):
        torch = self.torch
        if self.staged is not None:
            self._post()
        b = self.k & 1
        self.k += 1
        n = int(sim.num_hit())
        if n > self.bufs[b].shape[0]:
            # the transfer that last read this buffer was posted two events ago; wait for it before replacing the buffer
            self.stream.synchronize()
            self.bufs[b] = torch.empty((n + n // 4, 16), dtype=torch.float32, device=self.device)
        if n:
            sim.get_hits_device(self.bufs[b].data_ptr())        # async copy on the simulator's stream
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        # the next event may only overwrite this staging buffer's twin; this one is read by the gather posted at the next push
        self.staged = (b, n, ev)

    def drain(self):
        if self.staged is not None:
            self._post()
        self.stream.synchronize()
        return self.last

    def result(self):
        return self.last
