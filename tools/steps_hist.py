"""Distribution of executed roll-out steps per instance (how much of the lock-step batch is alive at each graph replay).
    python tools/steps_hist.py [batch]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import PARAMS, WEIGHTS, make_instances  # noqa: E402

from ctrm_b200 import RoadmapBuilder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = RoadmapBuilder(WEIGHTS, 0)
b.upload(make_instances(B, 0))
b.run(seed=2022, export=False, **PARAMS)
s = b.instance_steps().astype(np.int64)
tot = PARAMS["N_traj"] * PARAMS["max_T"]
print("replays", b.timing()["steps"], "steps/instance: min", s.min(), "mean", s.mean(), "median", np.median(s), "max", s.max(), "of", tot)
alive = np.array([(s > r).mean() for r in range(0, tot, 100)])
print("fraction of instances alive at replay 0,100,...:", np.round(alive, 3))
print("area under the alive curve / replays =", (np.minimum(s, tot).sum() / (B * float(b.timing()["steps"]))))
