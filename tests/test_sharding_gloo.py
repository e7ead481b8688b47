"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous instance sharding and the final gather of the
variable-length CSR results (ctrm_b200/parallel.py)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from ctrm_b200.parallel import gather_csr, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(7, rank, world)
    nv = 10 * (hi - lo) + rank  # ragged sizes across ranks
    res = dict(layer_ptr=torch.arange(nv + 1, dtype=torch.int32), pos=torch.full((nv, 2), float(rank), dtype=torch.float64),
               eptr=torch.arange(nv + 1, dtype=torch.int64) * 2, eidx=torch.full((2 * nv,), rank, dtype=torch.int32),
               n_layers=torch.full((hi - lo,), 66, dtype=torch.int32), cnt_collide=torch.arange(lo, hi, dtype=torch.int64))
    out = gather_csr(res, rank, world)
    # the asynchronous form used by bench.py (transfer overlaps the next step): same result once the works are waited for
    out2, works, keep = gather_csr(res, rank, world, async_op=True)
    for w in works:
        w.wait()
    del keep
    if rank == 0:
        ok = len(out) == world and len(out2) == world
        for r in range(world):
            for k in out[r]:
                ok &= bool(torch.equal(out[r][k], out2[r][k]))
        cover = []
        for r in range(world):
            l, h = shard_range(7, r, world)
            n = 10 * (h - l) + r
            ok &= out[r]["pos"].shape == (n, 2) and bool((out[r]["pos"] == float(r)).all())
            ok &= out[r]["eidx"].numel() == 2 * n and out[r]["layer_ptr"].numel() == n + 1
            cover += out[r]["cnt_collide"].tolist()
        ok &= cover == list(range(7))
        q.put(ok)
    else:
        assert out is None and out2 is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    from ctrm_b200.parallel import shard_range

    assert [shard_range(7, r, 2) for r in range(2)] == [(0, 4), (4, 7)]
    assert [shard_range(1024, r, 8) for r in range(8)][-1] == (896, 1024)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
