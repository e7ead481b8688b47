"""Generate tests/golden/* by RUNNING THE UNMODIFIED REFERENCE (build container only) and pin the oracle.

    python oracle/gen_golden.py            # writes fixtures, then checks the oracle against them

Fixtures (all small, committed):
  aamas22_main_best.npz   the trained networks' state_dicts + the hyper-parameters the path reads
                          (trained_models/aamas22-main/best_*)
  demo_ins.npz            notebooks/CTRM_demo_ins.pkl as plain arrays (config #1)
  ref_trace_demo.npz      one reference run (seed 0, N_traj=3) recorded through monkeypatches:
                          every np.random.rand draw in order, every np.sin/np.cos the random walk took,
                          the torch.rand Gumbel uniforms, per-step positions / X / SAMPLES, final TRMs
  ref_stage_*.npz         stage-level vectors (occupancy, cost-to-go, FOV, others_info, sweep verdicts)
                          produced by the reference's own functions on seeded inputs
Nothing here is imported by the product.
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

import ref_import  # noqa: E402

REF = ref_import.REF
BASENAME = os.path.join(REF, "trained_models", "aamas22-main", "best")


def write_weights():
    import torch

    arrays = {}
    for prefix, suffix in (("model", "_model.pt"), ("agent", "_agent_encoder.pt"),
                           ("fov_self", "_fov_encoder_self.pt"), ("fov_vain", "_fov_encoder_vain.pt")):
        sd = torch.load(BASENAME + suffix, map_location="cpu")
        for k, v in sd.items():
            if k.startswith("encoder_output") or k.endswith("num_batches_tracked"):
                continue  # train-only (cvae.py:66-68)
            arrays[f"{prefix}.{k}"] = v.numpy()
    hyp = pickle.load(open(BASENAME + "_model_hypra.pkl", "rb"))
    fov = pickle.load(open(BASENAME + "_fov_encoder_self.pkl", "rb"))
    age = pickle.load(open(BASENAME + "_agent_encoder.pkl", "rb"))
    fi = pickle.load(open(BASENAME + "_format_input.pkl", "rb"))  # needs the reference importable
    fo = pickle.load(open(BASENAME + "_format_output.Pl", "rb"))
    hp = dict(fov_size=fov["fov_size"], map_size=fov["map_size"], num_neighbors=fi.num_neighbors,
              use_k_neighbor=int(fi.use_k_neighbor), include_self_attention=int(fi.include_self_attention),
              dim_ind=fo.get_dim_indicators(), dim_latent=hyp["dim_latent"], dim_message=age["dim_message"],
              dim_attention=age["dim_attention"], temp=hyp["temp"])
    np.savez_compressed(os.path.join(GOLD, "aamas22_main_best.npz"), **arrays,
                        **{f"hp.{k}": np.array(v) for k, v in hp.items()})
    print("weights:", len(arrays), "arrays", hp)


def ins_to_arrays(ins):
    return dict(starts=np.array(ins.starts), goals=np.array(ins.goals), max_speeds=np.array(ins.max_speeds),
                rads=np.array(ins.rads), obs_pos=np.array([o.pos for o in ins.obs]).reshape(-1, 2),
                obs_rad=np.array([o.rad for o in ins.obs]))


def trms_to_arrays(trms, prefix=""):
    """flatten list[TimedRoadmap] (reference or oracle) -> arrays"""
    layer_ptr, pos, eptr, eidx = [0], [], [0], []
    nlayers = []
    for trm in trms:
        nlayers.append(len(trm.V))
        for t in range(len(trm.V)):
            for i, v in enumerate(trm.V[t]):
                pos.append(np.asarray(v.pos if hasattr(v, "pos") else v, dtype=np.float64))
                eidx.extend(int(e) for e in trm.E[t][i])
                eptr.append(len(eidx))
            layer_ptr.append(len(pos))
    return {prefix + "nlayers": np.array(nlayers, dtype=np.int32), prefix + "layer_ptr": np.array(layer_ptr, dtype=np.int32),
            prefix + "pos": np.array(pos, dtype=np.float64).reshape(-1, 2), prefix + "eptr": np.array(eptr, dtype=np.int32),
            prefix + "eidx": np.array(eidx, dtype=np.int32)}


class RefRecorder:
    """monkeypatches around ONE reference call; restores everything on exit"""

    def __init__(self, ctrm_mod):
        self.ctrm = ctrm_mod
        self.draws, self.sincos, self.uniforms, self.steps = [], [], [], []

    def __enter__(self):
        import torch
        import ctrm.roadmap_learned.learned_sampler as ls

        self._ls = ls
        self._rand, self._sin, self._cos, self._trand = np.random.rand, np.sin, np.cos, torch.rand
        self._reconstruct = ls.reconstruct
        rec = self

        def rand(*a):
            v = rec._rand(*a)
            if not a:
                rec.draws.append(float(v))
            return v

        def sin(x, *a, **k):
            v = rec._sin(x, *a, **k)
            if np.ndim(x) == 0:
                rec.sincos.append(float(v))
            return v

        def cos(x, *a, **k):
            v = rec._cos(x, *a, **k)
            if np.ndim(x) == 0:
                rec.sincos.append(float(v))
            return v

        def trand(*a, **k):
            v = rec._trand(*a, **k)
            rec.uniforms.append(v.numpy().copy())
            return v

        def reconstruct(basename):
            model, fi, fo = rec._reconstruct(basename)
            g = fi.get_feature_one_step
            s = model.sample

            def get_feature_one_step(ins, arr_current_pos, arr_prev_pos):
                X = g(ins=ins, arr_current_pos=arr_current_pos, arr_prev_pos=arr_prev_pos)
                rec.steps.append(dict(cur=np.array(arr_current_pos).copy(), prev=np.array(arr_prev_pos).copy(),
                                      X=X.numpy().copy()))
                return X

            def sample(x, ind):
                y = s(x, ind)
                rec.steps[-1]["IND"] = ind.numpy().copy()
                rec.steps[-1]["SAMPLES"] = y.numpy().copy()
                return y

            fi.get_feature_one_step = get_feature_one_step
            model.sample = sample
            rec.format_input = fi
            return model, fi, fo

        np.random.rand, np.sin, np.cos, torch.rand, ls.reconstruct = rand, sin, cos, trand, reconstruct
        return self

    def __exit__(self, *exc):
        import torch

        np.random.rand, np.sin, np.cos, torch.rand = self._rand, self._sin, self._cos, self._trand
        self._ls.reconstruct = self._reconstruct


def run_reference_trace(ins, seed, **kw):
    ctrm = ref_import.import_reference()
    from ctrm.roadmap_learned import get_timed_roadmaps_multiple_paths_with_learned_indicator as build
    from ctrm.utils import set_global_seeds

    set_global_seeds(seed)
    with RefRecorder(ctrm) as rec:
        trms = build(ins, BASENAME, **kw)
    return trms, rec


def write_demo_and_trace():
    ref_import.import_reference()
    ins = pickle.load(open(os.path.join(REF, "notebooks", "CTRM_demo_ins.pkl"), "rb"))
    np.savez_compressed(os.path.join(GOLD, "demo_ins.npz"), **ins_to_arrays(ins))
    kw = dict(N_traj=3, max_T=64, merge_distance_rate=0.1, max_attempt=3)
    trms, rec = run_reference_trace(ins, 0, **kw)
    S = len(rec.steps)
    out = dict(ins_to_arrays(ins))
    out.update(trms_to_arrays(trms, "trm."))
    out.update(seed=np.array(0), N_traj=np.array(kw["N_traj"]), max_T=np.array(kw["max_T"]),
               merge_distance_rate=np.array(kw["merge_distance_rate"]), max_attempt=np.array(kw["max_attempt"]),
               draws=np.array(rec.draws), sincos=np.array(rec.sincos),
               cur=np.stack([s["cur"] for s in rec.steps]), prev=np.stack([s["prev"] for s in rec.steps]),
               SAMPLES=np.stack([s["SAMPLES"] for s in rec.steps]),
               IND=np.stack([s["IND"] for s in rec.steps]).astype(np.int8),
               cnt_continuous_collide=np.array(ins.objs.cnt_continuous_collide))
    # stage-level NN goldens for a handful of steps only (keeps the fixture small)
    keep = sorted(set([0, 1, 2, S // 3, S // 2, S - 2, S - 1]))
    out.update(nn_steps=np.array(keep), nn_X=np.stack([rec.steps[i]["X"] for i in keep]),
               nn_U=np.stack([rec.uniforms[i] for i in keep]))
    np.savez_compressed(os.path.join(GOLD, "ref_trace_demo.npz"), **out)
    print(f"trace: {S} steps, {len(rec.draws)} draws, V={sum(len(v) for t in trms for v in t.V)}, "
          f"E={sum(len(e) for t in trms for l in t.E for e in l)}, checks={ins.objs.cnt_continuous_collide}")
    return ins, trms, rec, kw


# Reference runs with N_traj > dim_ind = 3, where the gate `u < p_rw` / `u > 0.9`, the p_rw(t, makespan) table and the
# dense-layer merge rules are DECIDED BY THE REFERENCE (learned_sampler.py:84-91, :148-168): with N_traj = 3 the clause
# `k_traj < dim_ind` short-circuits all of them.  One run with the eval parameters (merge_distance_rate 0.1,
# scripts/conf/eval/eval_large_ctrm_learned_ind.yaml) and one with the function default 0.25 (learned_sampler.py:29,
# notebooks/CTRM_demo.ipynb).  Only what the replay needs is kept: draws, sin/cos, SAMPLES, the final roadmaps, the counter.
EXTRA_TRACES = (("ref_trace_demo_nt8_r010.npz", 8, 0.1, 0), ("ref_trace_demo_nt6_r025.npz", 6, 0.25, 1))


def write_extra_traces(pin=True):
    ref_import.import_reference()
    for name, n_traj, rate, seed in EXTRA_TRACES:
        ins = pickle.load(open(os.path.join(REF, "notebooks", "CTRM_demo_ins.pkl"), "rb"))
        kw = dict(N_traj=n_traj, max_T=64, merge_distance_rate=rate, max_attempt=3)
        trms, rec = run_reference_trace(ins, seed, **kw)
        out = dict(ins_to_arrays(ins))
        out.update(trms_to_arrays(trms, "trm."))
        out.update(seed=np.array(seed), N_traj=np.array(n_traj), max_T=np.array(64), merge_distance_rate=np.array(rate),
                   max_attempt=np.array(3), draws=np.array(rec.draws), sincos=np.array(rec.sincos),
                   SAMPLES=np.stack([s["SAMPLES"] for s in rec.steps]),
                   cnt_continuous_collide=np.array(ins.objs.cnt_continuous_collide))
        np.savez_compressed(os.path.join(GOLD, name), **out)
        print(f"{name}: {len(rec.steps)} steps, {len(rec.draws)} draws, V={sum(len(v) for t in trms for v in t.V)}, "
              f"checks={ins.objs.cnt_continuous_collide}")
        if pin:
            pin_oracle(ins, trms, rec, kw)


def write_stage_goldens():
    """reference functions on seeded inputs: occupancy, cost-to-go, FOV, others_info, sweep verdicts"""
    ref_import.import_reference()
    import cost_to_go_wrapper
    import sphere_collision_check_wrapper as scw
    from ctrm.environment import ObstacleSphere, StaticObjects, generate_ins_2d_with_obs_hetero
    from ctrm.learning.compiled import get_arr_others_info
    from ctrm.learning.fov import get_approx_cost_to_go_matrix_2d, get_fov2d_occupancy_cost
    from ctrm.utils import set_global_seeds

    out = {}
    # (1) hetero instance (3 radii) on the trained map size: occupancy, cost-to-go, FOVs
    set_global_seeds(123)
    ins = generate_ins_2d_with_obs_hetero(num_agents=6, max_speeds_cands="0.03125,0.046875",
                                          rads_cands="0.015625,0.0234375,0.01", obs_num=12,
                                          obs_size_lower_bound=0.05, obs_size_upper_bound=0.12)
    M, Fv = 160, 19
    occ = ins.objs.draw_2d(M)
    ctg = [get_approx_cost_to_go_matrix_2d(ins.goals[i], occ, ins.rads[i]) for i in range(ins.num_agents)]
    rng = np.random.RandomState(7)
    pts = np.concatenate([rng.rand(20, 2), np.array([[0.0, 0.0], [1.0, 1.0], [0.999, 0.001], [0.0, 1.0], [0.5, 1.0]])])
    with np.errstate(invalid="ignore"):
        fovs = np.stack([np.stack([get_fov2d_occupancy_cost(p, occ, ctg[i], Fv, True) for p in pts])
                         for i in range(ins.num_agents)])
    ctg_i = np.stack([np.where(np.isinf(c), 65535, c).astype(np.int32) for c in ctg])
    out.update({f"s1.{k}": v for k, v in ins_to_arrays(ins).items()})
    out.update({"s1.occ": occ.astype(np.uint8), "s1.ctg": ctg_i.astype(np.uint16), "s1.pts": pts,
                "s1.fov": fovs.astype(np.uint8)})
    # (2) others_info (numba) on random positions, incl. coincident points (zero-norm branch)
    N = 7
    cur = rng.rand(N, 2)
    prev = rng.rand(N, 2)
    prev[2] = cur[2]
    goals = rng.rand(N, 2)
    goals[4] = cur[4]
    rads = rng.rand(N) * 0.02 + 0.01
    speeds = rng.rand(N) * 0.03 + 0.01
    info = get_arr_others_info(cur.reshape(1, N, 2), goals, prev.reshape(1, N, 2), rads, speeds).reshape(N * N, 11)
    out.update({"s2.cur": cur, "s2.prev": prev, "s2.goals": goals, "s2.rads": rads, "s2.speeds": speeds, "s2.info": info})
    # (3) swept-sphere verdicts from the reference's compiled helper, random + adversarial (tangent) cases
    n = 4000
    f1 = rng.rand(n, 2)
    t1 = f1 + (rng.rand(n, 2) - 0.5) * 0.08
    t1[:50] = f1[:50]  # degenerate from==to (point test)
    c = rng.rand(n, 2)
    r1 = rng.rand(n) * 0.03 + 0.005
    r2 = rng.rand(n) * 0.05 + 0.01
    # tangent: put the obstacle centre exactly (r1+r2) from the segment midpoint along the normal
    mid = 0.5 * (f1[100:600] + t1[100:600])
    d = t1[100:600] - f1[100:600]
    nrm = np.stack([-d[:, 1], d[:, 0]], 1) / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-300)
    c[100:600] = mid + nrm * (r1[100:600] + r2[100:600])[:, None]
    verdict = np.array([scw.continuousCollideSpheres(f1[i], t1[i], float(r1[i]), c[i], c[i], float(r2[i])) for i in range(n)])
    out.update({"s3.f1": f1, "s3.t1": t1, "s3.r1": r1, "s3.c": c, "s3.r2": r2, "s3.verdict": verdict})
    # (4) the reference's own test vectors, evaluated by the reference (tests/learning/test_fov.py:13-94)
    from ctrm.environment import Instance
    tins = Instance(num_agents=1, starts=[np.array([0.0, 0.0])], goals=[np.array([1.0, 1.0])], max_speeds=[0.1],
                    rads=[0.05], goal_rads=[0.01],
                    obs=[ObstacleSphere(np.array([0.5, 0.5]), 0.2), ObstacleSphere(np.array([1, 0]), 0.2)], dim=2)
    occ16 = tins.objs.draw_2d(16)
    ctg16 = get_approx_cost_to_go_matrix_2d(tins.goals[0], occ16, tins.rads[0])
    with np.errstate(invalid="ignore"):
        f00 = get_fov2d_occupancy_cost(np.array([0, 0]), occ16, ctg16, 5)
        f01 = get_fov2d_occupancy_cost(np.array([0, 1]), occ16, ctg16, 5)
    out.update({"s4.occ": occ16.astype(np.uint8), "s4.ctg": np.where(np.isinf(ctg16), 65535, ctg16).astype(np.uint16),
                "s4.fov00": f00.astype(np.uint8), "s4.fov01": f01.astype(np.uint8)})
    np.savez_compressed(os.path.join(GOLD, "ref_stage.npz"), **out)
    print("stage goldens written")


def pin_oracle(ins, trms, rec, kw):
    """the oracle, fed the reference's own streams, must reproduce the reference bit for bit"""
    import ctrm_oracle as O
    from rng_tables import ReplayRNG
    from ctrm.utils import set_global_seeds

    w = O.OWeights.load(os.path.join(GOLD, "aamas22_main_best.npz"))
    oins = O.OInstance.from_reference(ins)
    p = O.OParams(N_traj=kw["N_traj"], max_T=kw["max_T"], merge_distance_rate=kw["merge_distance_rate"],
                  max_attempt=kw["max_attempt"])
    recs = []
    otrms = O.build_timed_roadmaps(oins, w, p, ReplayRNG(rec.draws, rec.sincos, rec.uniforms), recorder=recs)
    a = trms_to_arrays(trms)
    b = trms_to_arrays(otrms)
    ok = all(np.array_equal(a[k], b[k]) for k in a)
    print("oracle(stream) == reference, bit for bit:", ok, "| steps", len(recs), "vs", len(rec.steps),
          "| collide count", oins.cnt_continuous_collide, "vs", ins.objs.cnt_continuous_collide)
    mx = max(float(np.abs(r["X"] - s["X"]).max()) for r, s in zip(recs, rec.steps))
    ms = max(float(np.abs(r["SAMPLES_nn"] - s["SAMPLES"]).max()) for r, s in zip(recs, rec.steps))
    print("max |X - X_ref| =", mx, " max |SAMPLES - SAMPLES_ref| =", ms)
    assert ok and oins.cnt_continuous_collide == ins.objs.cnt_continuous_collide


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    assert ref_import.available(), "run `make -C oracle ref` in the build container first"
    ref_import.import_reference()
    if "extra" in sys.argv[1:]:  # only the N_traj > dim_ind traces
        write_extra_traces()
        sys.exit(0)
    write_weights()
    write_stage_goldens()
    ins, trms, rec, kw = write_demo_and_trace()
    pin_oracle(ins, trms, rec, kw)
    write_extra_traces()
