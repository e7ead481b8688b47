"""Instance-parallel sharding over the GPUs of one box.

Timed-roadmap construction has no exchange step: instances are independent (the reference itself only parallelises
over instances, joblib in scripts/eval.py:108-111).  Each rank builds its shard with zero traffic; the only collective
is one final gather of the variable-length CSR results to rank 0 over NCCL / NVLink.
"""
from __future__ import annotations

from typing import Dict, List

def shard_range(n_instances: int, rank: int, world: int):
    """contiguous block partition: rank r owns [lo, hi); sizes differ by at most one"""
    base, rem = divmod(n_instances, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_csr(res: Dict[str, "torch.Tensor"], rank: int, world: int, dst: int = 0, async_op: bool = False):
    """gather every rank's CSR arrays (any dict of tensors with the same keys on every rank: the wide arrays of
    RoadmapBuilder.export_device or its compact byte-sized form) to `dst`.  Returns on dst a list (one dict per rank) of
    tensors, elsewhere None.  The element counts travel by one small all_gather; the payloads as EXACT-SIZE point-to-point
    transfers grouped into one batch (ncclGroupStart/End under NCCL: all peers stream into dst concurrently over
    NVLink) — no padding to the largest rank, no zero-fill, no staging copy.
    With async_op=True the transfers are only enqueued: returns (out, works, keepalive) and the caller calls `w.wait()` on
    every work before reading `out` (and keeps `keepalive` referenced until then), so that the traffic overlaps whatever the
    caller launches next."""
    import torch
    import torch.distributed as dist

    keys = sorted(res.keys())
    dev = res[keys[0]].device
    sizes = torch.tensor([res[k].numel() for k in keys], dtype=torch.int64, device=dev)
    all_sizes = torch.empty(world * len(keys), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes)
    out: List[dict] = [dict() for _ in range(world)] if rank == dst else None
    ops, keep = [], []
    if rank == dst:
        all_sizes = all_sizes.cpu().reshape(world, len(keys))  # only the receiver needs them on the host (to allocate)
        for r in range(world):
            for j, k in enumerate(keys):
                n = int(all_sizes[r, j])
                if r == dst:
                    buf = res[k].reshape(-1)
                else:
                    buf = torch.empty(n, dtype=res[k].dtype, device=dev)
                    if n:
                        ops.append(dist.P2POp(dist.irecv, buf, r))
                out[r][k] = buf.reshape(-1, 2) if res[k].dim() == 2 else buf
    else:
        for k in keys:
            flat = res[k].reshape(-1)
            if flat.numel():
                ops.append(dist.P2POp(dist.isend, flat, dst))
            keep.append(flat)
    works = dist.batch_isend_irecv(ops) if ops else []
    if async_op:
        return out, works, (keep, res)
    for w in works:
        w.wait()
    return out
