"""A/B of the two select+merge kernels (one warp per agent / one thread per agent) over batch sizes: per-launch time of the
`step` kernel from the eager profile pass and the whole build, same instances and seed.
    python tools/step_kernel_ab.py [sizes ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import PARAMS, WEIGHTS, make_instances  # noqa: E402

from ctrm_b200 import RoadmapBuilder  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [1, 16, 128, 512, 1024, 2048, 2960]
b = RoadmapBuilder(WEIGHTS, 0)
prm = dict(PARAMS, N_traj=6)
print("| instances | agents | step us (warp) | step us (thread) | build ms (warp) | build ms (thread) | kernels us/launch (thread run) |")
print("|---:|---:|---:|---:|---:|---:|---|")
for n in sizes:
    ins = make_instances(n, 0)
    packed = b.pack_instances(ins)
    row = []
    kt_last = None
    for mode in (1, 2):
        b.set_step_kernel(mode)
        b.upload(None, packed=packed)
        b.run(seed=1, export=False, **prm)
        b.upload(None, packed=packed)
        b.run(seed=1, export=False, **prm)
        build = b.timing()["rollout_ms"]
        b.set_profile(True)
        b.upload(None, packed=packed)
        b.run(seed=1, export=False, **prm)
        kt = b.kernel_times()
        b.set_profile(False)
        row.append((1e3 * kt["step"]["ms"] / max(1, kt["step"]["launches"]), build))
        kt_last = kt
    ks = " ".join(f"{k}={1e3 * v['ms'] / max(1, v['launches']):.1f}" for k, v in kt_last.items() if v["launches"])
    print(f"| {n} | {int(packed['n_agents'].sum())} | {row[0][0]:.1f} | {row[1][0]:.1f} | {row[0][1]:.1f} | {row[1][1]:.1f} | {ks} |")
b.close()
