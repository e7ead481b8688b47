"""Generate tests/golden/ref_pp_*.npz by RUNNING THE UNMODIFIED REFERENCE PLANNER (build container only) and pin
oracle/pp_oracle.py against it.

    python oracle/gen_golden_pp.py

Inputs are timed roadmaps of notebooks/CTRM_demo_ins.pkl: (a) the reference's own recorded roadmaps of
tests/golden/ref_trace_demo.npz (N_traj = 3), (b) roadmaps built by the CPU oracle with counter-based randomness
(N_traj = 25; committed with the fixture so that no test has to rebuild them).  Each is turned into the reference's
TimedRoadmap objects and solved by ctrm.planner.PrioritizedPlanning._solve (prioritized_planning.py:40-101); the fixture
holds the chosen vertex index per agent and time step and the planner's counters.  Nothing here is imported by the product.
"""
from __future__ import annotations

import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402

import ctrm_oracle as O  # noqa: E402
import pp_oracle as P  # noqa: E402
from rng_tables import PhiloxTables  # noqa: E402


def to_reference_trms(ctrm, arrays):
    from ctrm.roadmap import TimedNode, TimedRoadmap

    nlayers, layer_ptr, pos, eptr, eidx = arrays
    trms, k = [], 0
    for n in nlayers:
        n = int(n)
        trm = TimedRoadmap(np.array(pos[layer_ptr[k]], dtype=np.float64))
        trm.V, trm.E = [], []
        for t in range(n):
            v0, v1 = int(layer_ptr[k + t]), int(layer_ptr[k + t + 1])
            trm.V.append([TimedNode(t, i, np.array(pos[v0 + i], dtype=np.float64)) for i in range(v1 - v0)])
            trm.E.append([[int(j) for j in eidx[eptr[v]:eptr[v + 1]]] for v in range(v0, v1)])
        trms.append(trm)
        k += n
    return trms


def solve_reference(ctrm, ins, arrays):
    from ctrm.planner import PrioritizedPlanning

    pp = PrioritizedPlanning(ins, to_reference_trms(ctrm, arrays))
    np.random.seed(0)  # only breaks exact (f, t) ties
    pp._solve()
    paths = np.array([[v.index for v in path] for path in pp.solution], dtype=np.int32) if pp.solved else np.zeros((0, 0), np.int32)
    out = dict(solved=np.array(int(pp.solved)), paths=paths, cnt_continuous_collide=np.array(pp.cnt_continuous_collide),
               lowlevel_expanded=np.array(pp.lowlevel_expanded), lowlevel_explored=np.array(pp.lowlevel_explored))
    if pp.solved:
        assert pp.validate()
        out.update(sum_of_costs=np.array(pp.get_sum_of_costs(pp.solution)), maximum_costs=np.array(pp.get_maximum_costs(pp.solution)),
                   sum_of_travel_dists=np.array(pp.get_sum_of_travel_dists(pp.solution)))
    return out


def check_oracle(ins_arrays, arrays, ref):
    starts, goals, speeds, rads = ins_arrays
    got = P.prioritized_planning(starts, goals, speeds, rads, P.FlatRoadmaps(*arrays))
    assert got["solved"] == bool(ref["solved"]), (got["solved"], ref["solved"])
    for k in ("cnt_continuous_collide", "lowlevel_expanded", "lowlevel_explored"):
        assert got[k] == int(ref[k]), (k, got[k], int(ref[k]))
    if got["solved"]:
        assert np.array_equal(got["paths"], ref["paths"])
    print("  oracle == reference:", {k: int(ref[k]) for k in ("solved", "cnt_continuous_collide", "lowlevel_expanded",
                                                               "lowlevel_explored")})


def main():
    ctrm = ref_import.import_reference()
    ins = pickle.load(open(os.path.join(ref_import.REF, "notebooks", "CTRM_demo_ins.pkl"), "rb"))
    ins_arrays = (np.array(ins.starts), np.array(ins.goals), np.array(ins.max_speeds), np.array(ins.rads))
    out = {"goal_rads": np.array(ins.goal_rads, dtype=np.float64)}
    # (a) the reference's own roadmaps
    z = np.load(os.path.join(GOLD, "ref_trace_demo.npz"))
    arrays = tuple(z["trm." + k] for k in ("nlayers", "layer_ptr", "pos", "eptr", "eidx"))
    print("reference roadmaps (N_traj=3):", arrays[2].shape[0], "vertices")
    ref = solve_reference(ctrm, ins, arrays)
    check_oracle(ins_arrays, arrays, ref)
    out.update({"a." + k: v for k, v in ref.items()})
    # (b) oracle-built roadmaps, more roll-outs
    from helpers import otrms_to_arrays

    oi = O.OInstance.from_reference(ins)
    w = O.OWeights.load(os.path.join(GOLD, "aamas22_main_best.npz"))
    nt = int(os.environ.get("PP_GOLD_NTRAJ", "25"))
    trms = O.build_timed_roadmaps(oi, w, O.OParams(N_traj=nt, max_T=64, merge_distance_rate=0.1, max_attempt=3),
                                  PhiloxTables(11, 0))
    arrays = otrms_to_arrays(trms)
    print(f"oracle roadmaps (N_traj={nt}):", arrays[2].shape[0], "vertices,", arrays[4].shape[0], "edges")
    ref = solve_reference(ctrm, ins, arrays)
    check_oracle(ins_arrays, arrays, ref)
    out.update({"b." + k: v for k, v in ref.items()})
    out.update({"b.trm." + k: v for k, v in zip(("nlayers", "layer_ptr", "pos", "eptr", "eidx"), arrays)})
    out["b.trm.eidx"] = out["b.trm.eidx"].astype(np.int32)
    out["b.trm.layer_ptr"] = out["b.trm.layer_ptr"].astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, "ref_pp_demo.npz"), **out)
    print("wrote", os.path.join(GOLD, "ref_pp_demo.npz"), os.path.getsize(os.path.join(GOLD, "ref_pp_demo.npz")), "bytes")


if __name__ == "__main__":
    main()
