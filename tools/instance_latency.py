"""Latency of small batches: the per-instance cluster kernel (ctrm_set_step_kernel 3) against the batch path (1) at the
headline parameters (N_traj = 25, max_T = 64, K = 3).
    python tools/instance_latency.py [batch sizes, comma separated] [cluster sizes]
Prints device ms of the roll-out phase (CUDA events of ctrm_build_roadmaps), wall ms of upload + build + export, vertices.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import PARAMS, WEIGHTS, make_instances  # noqa: E402

from ctrm_b200 import RoadmapBuilder  # noqa: E402
from ctrm_b200.environment import Instance  # noqa: E402

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 8, 16, 32, 64, 128, 148, 296]
clusters = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
bld = RoadmapBuilder(WEIGHTS, 0)
z = np.load(os.path.join(ROOT, "tests", "golden", "demo_ins.npz"))
demo = Instance.from_arrays(z["starts"], z["goals"], z["max_speeds"], z["rads"], z["obs_pos"], z["obs_rad"])


def timed(instances, mode, reps=3):
    packed = bld.pack_instances(instances)
    bld.set_step_kernel(mode)
    best = None
    for r in range(reps + 1):
        t0 = time.perf_counter()
        bld.upload(None, packed=packed)
        csr = bld.run(seed=7, **PARAMS)
        wall = (time.perf_counter() - t0) * 1e3
        tim = bld.timing()
        if r > 0 and (best is None or tim["rollout_ms"] < best[0]):
            best = (tim["rollout_ms"], wall, csr.n_vertices, csr.n_edges)
    return best


print("| instances | kernel | cluster | roll-out ms (device) | wall ms | vertices | edges |")
print("|---|---|---|---|---|---|---|")
for B in sizes:
    ins = [demo] if B == 1 else make_instances(B, 424242)
    r = timed(ins, 1)
    print(f"| {B} | batch path | - | {r[0]:.2f} | {r[1]:.2f} | {r[2]} | {r[3]} |", flush=True)
    for C in clusters:
        if C:
            os.environ["CTRM_INSTANCE_CLUSTER"] = str(C)
        else:
            os.environ.pop("CTRM_INSTANCE_CLUSTER", None)
        r = timed(ins, 3)
        print(f"| {B} | per-instance | {C if C else 'auto'} | {r[0]:.2f} | {r[1]:.2f} | {r[2]} | {r[3]} |", flush=True)
bld.close()
