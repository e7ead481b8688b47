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
ins = make_instances(B, 0)
for depth in (1, 2, 3):
    pipe = PipelinedBuilder(WEIGHTS, 0, depth=depth)
    for _ in pipe.map([ins] * depth, seed=1, **PARAMS):
        pass
    t0 = time.perf_counter()
    nv = 0
    for csr in pipe.map([ins] * n, seed=1, **PARAMS):
        nv += csr.n_vertices
    dt = time.perf_counter() - t0
    print(f"depth {depth}: {1e3 * dt / n:.1f} ms/batch, {nv / dt:.3e} vertices/s; last phases (pack, upload, roll-outs, export) ms per context:",
          [tuple(round(x, 1) for x in b.last_phases_ms) for b in pipe.builders])
    pipe.close()
