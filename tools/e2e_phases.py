"""Where the end-to-end time of a stream of batches goes: host phases of RoadmapBuilder.build (pack, upload, roll-outs,
export) per batch and the throughput of PipelinedBuilder.map with 1, 2 and 3 contexts.
    python tools/e2e_phases.py [batch] [n_batches]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import PARAMS, WEIGHTS, make_instances  # noqa: E402

from ctrm_b200 import PipelinedBuilder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2960
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
depths = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 3]
dev = int(os.environ.get("LOCAL_RANK", "0"))  # under torchrun every rank drives its own GPU (no collectives: a host-side probe)
import torch  # noqa: E402

torch.cuda.set_device(dev)
ins = make_instances(B, dev * B)
for depth in depths:
    pipe = PipelinedBuilder(WEIGHTS, dev, depth=depth)
    for _ in pipe.map([ins] * depth, seed=1, **PARAMS):
        pass
    t0 = time.perf_counter()
    nv = 0
    for csr in pipe.map([ins] * n, seed=1, **PARAMS):
        nv += csr.n_vertices
    dt = time.perf_counter() - t0
    print(f"rank {dev} depth {depth}: {1e3 * dt / n:.1f} ms/batch, {nv / dt:.3e} vertices/s; last phases (pack, upload, roll-outs, export) ms per context:",
          [tuple(round(x, 1) for x in b.last_phases_ms) for b in pipe.builders])
    pipe.close()
