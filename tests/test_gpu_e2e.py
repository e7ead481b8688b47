"""End-to-end parity of the graph-captured GPU driver (ctrm_build_roadmaps) against
  (1) the REFERENCE's recorded run on notebooks/CTRM_demo_ins.pkl (tests/golden/ref_trace_demo.npz): fed the
      reference's own sampled coordinates (teacher forcing) and its own random draws (keyed tables), the GPU must
      reproduce the reference's timed roadmaps bit for bit — vertex positions, edge sets, layer structure and the
      collision-check counter;
  (2) the oracle run with counter-based randomness on a batch of two different instances: topology bit-exact under
      teacher forcing, and the GPU's own CVAE outputs within 1e-5 relative of the oracle's at EVERY executed step."""
import numpy as np
import pytest

import ctrm_oracle as O
from helpers import (KeyedReplay, assert_same_roadmaps, csr_to_arrays, golden_trm_arrays, instance_from_oins, oins_from_npz,
                     otrms_to_arrays, rel_err, teacher_table)
from rng_tables import PhiloxTables

pytestmark = pytest.mark.gpu


def test_reference_run_replay_bit_exact(builder, any_trace_gold, oracle_weights):
    z = any_trace_gold
    oi = oins_from_npz(z)
    N, K = oi.num_agents, 3
    NT, MT, MA = int(z["N_traj"]), int(z["max_T"]), int(z["max_attempt"])
    rate = float(z["merge_distance_rate"])
    p = O.OParams(N_traj=NT, max_T=MT, merge_distance_rate=rate, max_attempt=MA)
    # map the stream-order draws of the reference onto keys by replaying them through the oracle, teacher-forced with
    # the reference's SAMPLES in step order
    steps = iter(range(len(z["SAMPLES"])))

    class StepTeacher(dict):
        def __getitem__(self, key):
            return z["SAMPLES"][next(steps)]

    rng = KeyedReplay(z["draws"], z["sincos"], None, N, NT, MT, MA)
    recs = []
    otrms = O.build_timed_roadmaps(oi, oracle_weights, p, rng, recorder=recs, teacher=StepTeacher())
    want = golden_trm_arrays(z)
    assert_same_roadmaps(otrms_to_arrays(otrms), want)  # the oracle replay itself reproduces the reference
    teacher = teacher_table(recs, N, K, NT, MT)
    gate = np.nan_to_num(rng.t_gate, nan=0.5)
    rw = np.nan_to_num(rng.t_rw, nan=0.0)
    builder.upload([instance_from_oins(oi)])
    for mode in (2, 1, 3, 0):  # thread-per-agent kernel, warp-per-agent kernel, per-instance kernel, then the default choice
        builder.set_step_kernel(mode)
        csr = builder.run(N_traj=NT, max_T=MT, max_attempt=MA, merge_distance_rate=rate,
                          tables=dict(gate=gate, rw=rw, teacher=teacher))
        assert_same_roadmaps(csr_to_arrays(csr, 0), want)
        assert int(csr.cnt_collide[0]) == int(z["cnt_continuous_collide"])
    nv, ne = csr.instance_counts(0)
    assert (nv, ne) == (len(want[2]), len(want[4]))
    # planner contract (prioritized_planning.py:41, :57-59): equal lengths, V[0] = [start], V[t][-1] = goal
    trms = csr.timed_roadmaps(0)
    assert len({len(t.V) for t in trms}) == 1
    for i, trm in enumerate(trms[:3]):
        assert np.array_equal(trm.V[0][0].pos, oi.starts[i])
        assert all(np.array_equal(trm.V[t][-1].pos, oi.goals[i]) for t in range(1, len(trm.V)))


@pytest.mark.parametrize("rate", [0.1, 0.25])
def test_philox_batch_teacher_forced(builder, trace_gold, stage_gold, oracle_weights, rate):
    oins = [oins_from_npz(trace_gold), oins_from_npz(stage_gold, "s1.")]
    NT, MT, MA, K, seed = 2, 64, 3, 3, 20221019
    p = O.OParams(N_traj=NT, max_T=MT, merge_distance_rate=rate, max_attempt=MA)
    all_recs, all_trms = [], []
    for b, oi in enumerate(oins):
        recs = []
        all_trms.append(O.build_timed_roadmaps(oi, oracle_weights, p, PhiloxTables(seed, 7 + b), recorder=recs))
        all_recs.append(recs)
    teacher = np.concatenate([teacher_table(r, oi.num_agents, K, NT, MT, key="SAMPLES_nn")
                              for r, oi in zip(all_recs, oins)], axis=2)
    builder.upload([instance_from_oins(oi) for oi in oins])
    csr = builder.run(N_traj=NT, max_T=MT, max_attempt=MA, merge_distance_rate=rate, seed=seed, instance_id_base=7,
                      tables=dict(teacher=teacher), trace=True)
    for b, oi in enumerate(oins):
        assert_same_roadmaps(csr_to_arrays(csr, b), otrms_to_arrays(all_trms[b]))
        assert int(csr.cnt_collide[b]) == oi.cnt_continuous_collide
    # stage-wise parity along the whole trajectory
    samples, loc_pred, loc_next = builder.read_trace()
    a0 = 0
    worst = 0.0
    for recs, oi in zip(all_recs, oins):
        N = oi.num_agents
        executed = np.zeros((NT, MT), dtype=bool)
        for r in recs:
            k, t = r["k_traj"], r["t"] - 1
            executed[k, t] = True
            assert np.array_equal(loc_pred[k, t, a0:a0 + N], np.array(r["loc_pred"]))
            assert np.array_equal(loc_next[k, t, a0:a0 + N], np.array(r["loc_next"]))
            ref = r["SAMPLES_nn"].reshape(K, N, 3).transpose(1, 0, 2)
            worst = max(worst, rel_err(samples[k, t, a0:a0 + N], ref))
        assert np.isnan(loc_next[~executed][:, a0:a0 + N]).all()  # steps after an early finish never ran
        a0 += N
    assert worst <= 1e-5, worst


def test_free_running_is_statistically_close(builder, trace_gold, oracle_weights):
    """without teacher forcing fp32 summation order can flip a discrete decision (the reference itself differs between
    1 and 8 torch threads): sizes must agree closely, not bit for bit"""
    oi = oins_from_npz(trace_gold)
    NT, seed = 3, 99
    p = O.OParams(N_traj=NT, max_T=64, merge_distance_rate=0.1, max_attempt=3)
    otrms = O.build_timed_roadmaps(oi, oracle_weights, p, PhiloxTables(seed, 0))
    ov, oe = O.trm_counts(otrms)
    builder.upload([instance_from_oins(oi)])
    csr = builder.run(N_traj=NT, max_T=64, max_attempt=3, merge_distance_rate=0.1, seed=seed)
    assert abs(csr.n_vertices - ov) <= 0.03 * ov, (csr.n_vertices, ov)
    assert abs(csr.n_edges - oe) <= 0.08 * oe, (csr.n_edges, oe)


def test_drop_in_function_and_planner_contract(trace_gold):
    from conftest import WEIGHTS
    from ctrm_b200 import get_timed_roadmaps_multiple_paths_with_learned_indicator as build

    ins = instance_from_oins(oins_from_npz(trace_gold))
    trms = build(ins, WEIGHTS, N_traj=2, merge_distance_rate=0.1, seed=3)
    assert len(trms) == ins.num_agents
    assert len({len(t.V) for t in trms}) == 1
    assert ins.objs.cnt_continuous_collide > 0
    for i, trm in enumerate(trms):
        assert len(trm.V[0]) == 1 and np.array_equal(trm.V[0][0].pos, ins.starts[i])
        for t in range(1, len(trm.V)):
            assert np.array_equal(trm.V[t][-1].pos, ins.goals[i])
            assert len(trm.E[t]) == len(trm.V[t])
            if t + 1 < len(trm.V):
                assert all(0 <= j < len(trm.V[t + 1]) for e in trm.E[t] for j in e)
    again = build(ins, WEIGHTS, N_traj=2, merge_distance_rate=0.1, seed=3)
    assert all(np.array_equal(a.getNodesPosOnestep(5), b.getNodesPosOnestep(5)) for a, b in zip(trms, again))
