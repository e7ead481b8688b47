"""Golden vectors for the baseline samplers, written by RUNNING THE UNMODIFIED REFERENCE in the
build container:

    python oracle/gen_golden_samplers.py        # -> tests/golden/ref_samplers.npz, then pins oracle/samplers_oracle.py

The reference's samplers query FCL once (StaticObjects.collide_sphere); FCL is absent, so oracle/ref_stubs/fcl.py answers that
query with FCL 0.5.0's published sphere-sphere test (parity unpinned at that boundary).  Everything else — sampling order and
RNG consumption, valid_move, adjacency order, append_fixed_strucutre, the two counters — is the reference's own code.
Cases: the instance of the reference's tests/roadmap/test_random_sampler.py:13-25 / test_grid_sampler.py (2 agents, speed 0.5)
and a seeded aamas22-shape instance (8 agents, 10 obstacles, speed 1/32).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)

import ref_import  # noqa: E402


def flatten_ref(trms):
    pos, eptr, eidx = [], [0], []
    for trm in trms:
        for t in range(len(trm.V)):
            for i, v in enumerate(trm.V[t]):
                pos.append(np.asarray(v.pos, dtype=np.float64))
                eidx.extend(int(e) for e in trm.E[t][i])
                eptr.append(len(eidx))
    return np.array(pos).reshape(-1, 2), np.array(eptr, dtype=np.int64), np.array(eidx, dtype=np.int32)


CASES = (  # name, instance, sampler, kwargs, seed
    ("t_random", "test", "random", dict(T=3, num=10), 1),
    ("t_fully", "test", "fully_random", dict(T=3, num=10), 2),
    ("t_rcommon", "test", "random_common", dict(T=3, num=10), 3),
    ("t_grid", "test", "grid", dict(T=3, size=5), 4),
    ("t_gcommon", "test", "grid_common", dict(T=3, size=5), 5),
    ("a_random", "aamas", "random", dict(T=4, num=300), 6),
    ("a_random_inv", "aamas", "random", dict(T=2, num=200, wo_invalid_samples=False), 7),
    ("a_rcommon", "aamas", "random_common", dict(T=3, num=300), 8),
    ("a_grid", "aamas", "grid", dict(T=3, size=24), 9),
    ("a_gcommon", "aamas", "grid_common", dict(T=3, size=40), 10),
    ("a_fully", "aamas", "fully_random", dict(T=3, num=60), 11),
)


def instances():
    ref_import.import_reference()
    from ctrm.environment import Instance, ObstacleSphere, generate_ins_2d_with_obs_hetero
    from ctrm.utils import set_global_seeds

    test = Instance(2, [np.array([0.0, 0.0]), np.array([1.0, 0.0])], [np.array([1.0, 1.0]), np.array([0.0, 1.0])], [0.5, 0.5],
                    [0.1, 0.1], [0.1, 0.1], [ObstacleSphere(pos=np.array([0.5, 0.5]), rad=0.2)], 2)
    set_global_seeds(321)
    aamas = generate_ins_2d_with_obs_hetero(num_agents=8, max_speeds_cands="0.03125", rads_cands="0.015625", obs_num=10,
                                            obs_size_lower_bound=0.05, obs_size_upper_bound=0.08)
    return dict(test=test, aamas=aamas)


def main():
    ref_import.import_reference()
    import ctrm.roadmap as RM
    from gen_golden import ins_to_arrays

    fns = dict(random=RM.get_timed_roadmaps_random, fully_random=RM.get_timed_roadmaps_fully_random,
               random_common=RM.get_timed_roadmaps_random_common, grid=RM.get_timed_roadmaps_grid,
               grid_common=RM.get_timed_roadmaps_grid_common)
    out = {}
    inss = instances()
    for key, ins in inss.items():
        out.update({f"{key}.{k}": v for k, v in ins_to_arrays(ins).items()})
    for name, which, sampler, kw, seed in CASES:
        ins = inss[which]
        ins.objs.cnt_static_collide = 0
        ins.objs.cnt_continuous_collide = 0
        np.random.seed(seed)
        trms = fns[sampler](ins, **kw)
        pos, eptr, eidx = flatten_ref(trms)
        after = np.random.random()  # where the global stream stands after the call
        out.update({f"{name}.pos": pos, f"{name}.eptr": eptr, f"{name}.eidx": eidx,
                    f"{name}.nlayers": np.array([len(t.V) for t in trms]),
                    f"{name}.layer_sizes": np.array([len(l) for t in trms for l in t.V]),
                    f"{name}.cnt_static": np.array(ins.objs.cnt_static_collide),
                    f"{name}.cnt_continuous": np.array(ins.objs.cnt_continuous_collide), f"{name}.rng_after": np.array(after)})
        print(name, "V", len(pos), "E", len(eidx), "static", ins.objs.cnt_static_collide, "cont", ins.objs.cnt_continuous_collide)
    np.savez_compressed(os.path.join(GOLD, "ref_samplers.npz"), **out)
    pin(out)


def pin(z):
    """oracle/samplers_oracle.py, run on the same seeds, must reproduce the reference bit for bit"""
    import ctrm_oracle as O
    import samplers_oracle as S

    for name, which, sampler, kw, seed in CASES:
        oi = O.OInstance.from_arrays(z[f"{which}.starts"], z[f"{which}.goals"], z[f"{which}.max_speeds"], z[f"{which}.rads"],
                                     z[f"{which}.obs_pos"], z[f"{which}.obs_rad"])
        cnt = S.Counters()
        np.random.seed(seed)
        fn = dict(random=S.random, fully_random=S.fully_random, random_common=S.random_common, grid=S.grid,
                  grid_common=S.grid_common_2d_fast)[sampler]
        trms = fn(oi, cnt, **kw)
        pos, eptr, eidx = S.flatten(trms)
        ok = (np.array_equal(pos, z[f"{name}.pos"]) and np.array_equal(eptr, z[f"{name}.eptr"]) and
              np.array_equal(eidx, z[f"{name}.eidx"]) and cnt.cnt_static_collide == int(z[f"{name}.cnt_static"]) and
              oi.cnt_continuous_collide == int(z[f"{name}.cnt_continuous"]) and np.random.random() == float(z[f"{name}.rng_after"]))
        print("oracle == reference:", name, ok)
        assert ok, name


if __name__ == "__main__":
    main()
