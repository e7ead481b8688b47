"""Instance-parallel sharding over the GPUs of one box.

Timed-roadmap construction has no exchange step: instances are independent (the reference itself only parallelises
over instances, joblib in scripts/eval.py:108-111).  Each rank builds its shard with zero traffic; the only collective
is one final gather of the variable-length CSR results to rank 0 over NCCL / NVLink.
"""
from __future__ import annotations

from typing import Dict, List

KEYS = ("layer_ptr", "pos", "eptr", "eidx", "n_layers", "cnt_collide")


def shard_range(n_instances: int, rank: int, world: int):
    """contiguous block partition: rank r owns [lo, hi); sizes differ by at most one"""
    base, rem = divmod(n_instances, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_csr(res: Dict[str, "torch.Tensor"], rank: int, world: int, dst: int = 0, async_op: bool = False):
    """gather every rank's CSR arrays to `dst`.  Returns on dst a list (one dict per rank) of trimmed tensors, elsewhere
    None.  Sizes travel by all_gather, payloads by a padded gather (equal-size requirement of the collective).
    With async_op=True the payload gathers are only enqueued: returns (out, works, keepalive) and the caller calls
    `w.wait()` on every work before reading `out` (and keeps `keepalive` referenced until then), so that the transfer over
    NVLink overlaps whatever the caller launches next."""
    import torch
    import torch.distributed as dist

    dev = res["pos"].device
    sizes = torch.tensor([res[k].numel() for k in KEYS], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).cpu()
    out: List[dict] = [dict() for _ in range(world)] if rank == dst else None
    works, keep = [], []
    for j, k in enumerate(KEYS):
        mx = int(all_sizes[:, j].max())
        flat = res[k].reshape(-1)
        pad = torch.zeros(mx, dtype=flat.dtype, device=dev)
        pad[:flat.numel()] = flat
        bufs = [torch.empty(mx, dtype=flat.dtype, device=dev) for _ in range(world)] if rank == dst else None
        if async_op:
            works.append(dist.gather(pad, bufs, dst=dst, async_op=True))
            keep.append((pad, bufs))
        else:
            dist.gather(pad, bufs, dst=dst)
        if rank == dst:
            for r in range(world):
                t = bufs[r][:int(all_sizes[r, j])]
                out[r][k] = t.reshape(-1, 2) if k == "pos" else t
    return (out, works, keep) if async_op else out
