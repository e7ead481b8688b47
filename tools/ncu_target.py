"""Short single-GPU run of the hot path for ncu: one batch of synthetic aamas22 instances, a few roll-outs.
    python tools/ncu_target.py [batch] [n_traj] [plan|eager]
`eager` launches the step kernels one by one (the context's profile mode) instead of replaying 32-step CUDA graphs, so that
`ncu -s <skip> -c <count>` can address single launches; with n_traj >= 4 the launches past ~1,300 belong to a roll-out whose
gate is active (k_traj >= 3): the network kernels then run over the learned rows only, as in the benchmark's steady state.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import PARAMS, WEIGHTS, make_instances  # noqa: E402

from ctrm_b200 import RoadmapBuilder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 2
b = RoadmapBuilder(WEIGHTS, 0)
if len(sys.argv) > 3 and sys.argv[3] == "eager":
    b.set_profile(True)
prm = dict(PARAMS, N_traj=NT)
csr = b.build(make_instances(B, 0), seed=1, **prm)
print("vertices", csr.n_vertices, "edges", csr.n_edges, csr.timing)
if len(sys.argv) > 3 and sys.argv[3] == "plan":  # also run the planner on the device-resident roadmaps
    r = b.plan()
    print("planner ms", r["elapsed_ms"], "solved", float(r["solved"].float().mean()))
