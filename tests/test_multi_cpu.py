"""world_size-2 gloo test of the N>1 path's host logic: event sharding with absolute photon offsets
and the variable-length hit gather (counts, then padded records) in rank order."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from eic_opticks_b200 import parallel, workloads
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = workloads.sipm8x8_scint(num_photon=20000, photons_per_genstep=100)
    gs_r, ip_r, off, cnt = parallel.shard_event(w["gensteps"], rank, world)
    # stand-in for the simulation: every 7th photon of this rank's absolute range is a "hit" carrying its index
    idx = np.arange(off, off + cnt, dtype=np.uint32)
    sel = idx[idx % 7 == rank % 2]
    hits = np.zeros((len(sel), 4, 4), dtype=np.float32)
    hits.view(np.uint32)[:, 3, 2] = sel
    allhits, counts = parallel.gather_hits(hits)
    # root-only gather (the reference hands the hits to one host process): same array on the root, None elsewhere
    root, counts_r = parallel.gather_hits(hits, dst=1)
    assert counts_r == counts
    if rank == 1:
        assert root.shape == allhits.shape and root.tobytes() == allhits.tobytes()
    else:
        assert root is None
    empty, counts_e = parallel.gather_hits(hits[:0] if rank == 0 else hits, dst=0)     # a rank without hits
    if rank == 0:
        assert len(empty) == counts[1] and counts_e[0] == 0
    q.put((rank, off, cnt, counts, allhits.view(np.uint32)[:, 3, 2].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, off0, cnt0, counts0, all0), (r1, off1, cnt1, counts1, all1) = res
    assert off0 == 0 and off1 == cnt0 and cnt0 + cnt1 == 20000
    assert counts0 == counts1 and sum(counts0) == len(all0)
    assert (all0 == all1).all()
    assert (np.diff(all0.astype(np.int64)) > 0).all()                      # rank order == ascending photon index
    assert (all0[:counts0[0]] < cnt0).all() and (all0[counts0[0]:] >= cnt0).all()
