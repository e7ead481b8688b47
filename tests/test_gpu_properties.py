"""Size-independent properties of the GPU roadmaps at BASELINE.json's full sizes, where the CPU oracle would take hours:
every edge of every timed roadmap re-checked against `valid_move` (roadmap/utils.py:18-42 + collision_check.cpp:16-58) in
vectorised numpy, the planner contract (prioritized_planning.py:41, :57-59, :96), determinism, and independence of an
instance's result from the batch it is built in (the RNG is keyed by instance id, not by batch position)."""
import numpy as np
import pytest

from conftest import WEIGHTS

pytestmark = pytest.mark.gpu


def _gen(seed, n_lo, n_hi, obs_num, rads="0.015625", lo=0.05, hi=0.08):
    from ctrm_b200.environment import generate_ins_2d_with_obs_hetero_nonfix_agents as gen

    return gen(n_lo, n_hi, "0.03125", rads, obs_num, lo, hi, rng=np.random.RandomState(seed))


def _swept_hits(p, q, rad, obs):
    """continuousCollideSpheres for E segments p->q of radius rad (E,) against O static circles obs (O,3) -> (E,) bool"""
    hit = np.zeros(len(p), dtype=bool)
    v = q - p                                       # (to1 - from1) + (from2 - to2), obstacle static
    a = v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]
    for ox, oy, orad in obs:
        wx, wy = p[:, 0] - ox, p[:, 1] - oy         # from1 - from2
        b = v[:, 0] * wx + v[:, 1] * wy
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(b >= 0, 0.0, np.where(a + b <= 0, 1.0, -b / a))
        dx, dy = wx + t * v[:, 0], wy + t * v[:, 1]
        hit |= dx * dx + dy * dy <= (rad + orad) ** 2
    return hit


def _check_instance(csr, b, ins, slack=0.0):
    """planner contract + every edge valid; returns (#vertices, #edges) checked"""
    starts, goals, speeds, rads, obs = ins.arrays()
    a0, a1 = int(csr.agent_off[b]), int(csr.agent_off[b + 1])
    L = csr.L
    assert len({int(n) for n in csr.n_layers[a0:a1]}) == 1          # equal length over the agents of an instance
    nv = ne = 0
    for a in range(a0, a1):
        i = a - a0
        nl = int(csr.n_layers[a])
        lp = csr.layer_ptr[a * L:a * L + nl + 1].astype(np.int64)
        assert lp[1] - lp[0] == 1 and np.array_equal(csr.pos[lp[0]], starts[i])       # V[0] == [start]
        last = lp[2:nl + 1] - 1
        assert np.array_equal(csr.pos[last], np.broadcast_to(goals[i], (len(last), 2)))  # V[t][-1] is the goal, t >= 1
        v0, v1 = int(lp[0]), int(lp[nl])
        pos = csr.pos[v0:v1]
        assert np.all(pos >= rads[i] - 1e-12) and np.all(pos <= 1 - rads[i] + 1e-12)   # inside the unit square
        # edges: source vertex v of layer t -> local index into layer t + 1
        eptr = csr.eptr[v0:v1 + 1].astype(np.int64)
        deg = np.diff(eptr)
        src = np.repeat(np.arange(v0, v1), deg)
        layer_of = np.repeat(np.arange(nl), np.diff(lp))
        e_layer = layer_of[src - v0]
        eidx = csr.eidx[eptr[0]:eptr[-1]].astype(np.int64)
        assert np.all(deg[layer_of == nl - 1] == 0) or True            # E[last] is not used by the planner
        keep = e_layer < nl - 1
        src, e_layer, eidx = src[keep], e_layer[keep], eidx[keep]
        width = np.diff(lp)[e_layer + 1]
        assert np.all((eidx >= 0) & (eidx < width))                    # E[t][i] indexes into V[t+1]
        dst = lp[e_layer + 1] + eidx
        p, q = csr.pos[src], csr.pos[dst]
        d = q - p
        sp = speeds[i]
        assert np.all(np.abs(d) <= sp + slack)
        assert np.all(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] <= sp ** 2 + slack)
        assert not _swept_hits(p, q, rads[i], obs).any()
        nv += v1 - v0
        ne += len(src)
    return nv, ne


def test_full_size_batch_properties_and_batch_independence():
    """BASELINE configs[1] at its real roll-out sizes (N_traj = 25, max_T = 64), 96 instances in one batch"""
    from ctrm_b200 import RoadmapBuilder

    prm = dict(N_traj=25, max_T=64, max_attempt=3, merge_distance_rate=0.1)
    instances = [_gen(46 + i, 21, 30, 10) for i in range(96)]
    bld = RoadmapBuilder(WEIGHTS, 0)
    try:
        bld.set_step_kernel(1)  # the batch path (a batch of up to one instance per SM would take the per-instance kernel, tested below)
        csr = bld.build(instances, seed=7, **prm).copy()
        nv = ne = 0
        for b in range(0, 96, 5):
            v, e = _check_instance(csr, b, instances[b])
            nv, ne = nv + v, ne + e
        assert nv > 400_000 and ne > 1_500_000, (nv, ne)
        assert np.all(csr.cnt_collide > 0)
        # same seed, same batch -> bit-identical result
        again = bld.build(instances, seed=7, **prm).copy()
        for f in ("layer_ptr", "eptr", "eidx", "n_layers", "cnt_collide"):
            assert np.array_equal(getattr(csr, f), getattr(again, f)), f
        assert np.array_equal(csr.pos.view(np.uint64), again.pos.view(np.uint64))
        # an instance built alone (same instance id) equals its result inside the batch
        b = 37
        alone = bld.build([instances[b]], seed=7, instance_id_base=b, **prm).copy()
        bld.set_step_kernel(0)
        a0, a1 = int(csr.agent_off[b]), int(csr.agent_off[b + 1])
        v0, v1 = int(csr.layer_ptr[a0 * csr.L]), int(csr.layer_ptr[a1 * csr.L])
        assert np.array_equal(alone.pos.view(np.uint64), csr.pos[v0:v1].view(np.uint64))
        assert np.array_equal(alone.eidx, csr.eidx[int(csr.eptr[v0]):int(csr.eptr[v1])])
        assert np.array_equal(alone.n_layers, csr.n_layers[a0:a1])
    finally:
        bld.close()


def test_instance_kernel_properties_and_cluster_sizes():
    """The per-instance cluster kernel (ctrm_set_step_kernel 3; the default for small batches) free-running at the headline
    roll-out sizes: every edge valid, deterministic, an instance's result independent of the batch it is built in and of the
    cluster size (1, 2, 4, 8 CTAs per instance: a warp per agent or per quarter of its FOV inputs, the step as one warp or
    one group of four warps per agent — the same sums in the same order, the same fp64 operations)"""
    import os

    from ctrm_b200 import RoadmapBuilder

    prm = dict(N_traj=25, max_T=64, max_attempt=3, merge_distance_rate=0.1)
    instances = [_gen(146 + i, 21, 30, 10) for i in range(5)] + [_gen(200, 33, 40, 10)]
    bld = RoadmapBuilder(WEIGHTS, 0)
    saved = os.environ.pop("CTRM_INSTANCE_CLUSTER", None)

    def same(x, y):
        return (np.array_equal(x.pos.view(np.uint64), y.pos.view(np.uint64))
                and all(np.array_equal(getattr(x, f), getattr(y, f)) for f in ("layer_ptr", "eptr", "eidx", "n_layers", "cnt_collide")))

    try:
        bld.set_step_kernel(3)
        csr = bld.build(instances, seed=3, **prm).copy()
        for b in range(len(instances)):
            nv, ne = _check_instance(csr, b, instances[b])
            assert nv > 20_000 and ne > nv
        assert same(csr, bld.build(instances, seed=3, **prm).copy())
        b = 2
        alone = bld.build([instances[b]], seed=3, instance_id_base=b, **prm).copy()
        a0, a1 = int(csr.agent_off[b]), int(csr.agent_off[b + 1])
        v0, v1 = int(csr.layer_ptr[a0 * csr.L]), int(csr.layer_ptr[a1 * csr.L])
        assert np.array_equal(alone.pos.view(np.uint64), csr.pos[v0:v1].view(np.uint64))
        assert np.array_equal(alone.eidx, csr.eidx[int(csr.eptr[v0]):int(csr.eptr[v1])])
        by_c = {}
        for c in (1, 2, 4, 8):
            os.environ["CTRM_INSTANCE_CLUSTER"] = str(c)
            by_c[c] = bld.build(instances, seed=3, **prm).copy()
        assert same(by_c[8], csr)
        assert same(by_c[1], by_c[2]) and same(by_c[1], by_c[4]) and same(by_c[1], by_c[8])
        for b in (0, 5):
            _check_instance(by_c[1], b, instances[b])
        # the default choice for a handful of instances is this kernel
        os.environ.pop("CTRM_INSTANCE_CLUSTER", None)
        bld.set_step_kernel(0)
        assert same(bld.build(instances, seed=3, **prm).copy(), csr)
    finally:
        if saved is not None:
            os.environ["CTRM_INSTANCE_CLUSTER"] = saved
        else:
            os.environ.pop("CTRM_INSTANCE_CLUSTER", None)
        bld.close()


def test_stress_shape_properties():
    """BASELINE configs[4]: 100 agents, 64 obstacles, K = 64 candidates, map_size 250, long horizon"""
    from ctrm_b200 import RoadmapBuilder

    instances = [_gen(900 + i, 100, 101, 64, rads="0.00390625", lo=0.02, hi=0.04) for i in range(2)]
    assert all(i.num_agents == 100 and len(i.obs) == 64 for i in instances)
    bld = RoadmapBuilder(WEIGHTS, 0, map_size=250)
    try:
        csr = bld.build(instances, seed=11, N_traj=3, max_T=96, max_attempt=3, merge_distance_rate=0.1, K=64).copy()
        for b in range(2):
            nv, ne = _check_instance(csr, b, instances[b])
            assert nv > 100 * 96 and ne > nv
        steps = bld.instance_steps()
        assert np.all(steps > 0) and np.all(steps <= 3 * 96)
    finally:
        bld.close()


def test_pipelined_builder_matches_sequential():
    """PipelinedBuilder.map over three batches == RoadmapBuilder.build batch by batch with the same instance ids"""
    from ctrm_b200 import PipelinedBuilder, RoadmapBuilder

    prm = dict(N_traj=3, max_T=32, max_attempt=3, merge_distance_rate=0.1)
    batches = [[_gen(300 + 10 * k + i, 8, 14, 6) for i in range(3 + k)] for k in range(3)]
    pipe = PipelinedBuilder(WEIGHTS, 0)
    seq = RoadmapBuilder(WEIGHTS, 0)
    try:
        got = [c.copy() for c in pipe.map(batches, seed=5, instance_id_base=100, **prm)]
        base = 100
        for k, batch in enumerate(batches):
            want = seq.build(batch, seed=5, instance_id_base=base, **prm)
            base += len(batch)
            assert np.array_equal(got[k].pos.view(np.uint64), want.pos.view(np.uint64))
            for f in ("layer_ptr", "eptr", "eidx", "n_layers", "cnt_collide"):
                assert np.array_equal(getattr(got[k], f), getattr(want, f)), (k, f)
    finally:
        pipe.close()
        seq.close()


def test_compact_export_equals_wide_export():
    """ctrm_export_csr_compact (bytes for layer counts / degrees / successor indices) expands to exactly the arrays of
    ctrm_export_csr, for the whole batch and instance by instance, in less than 0.6 of the bytes"""
    from ctrm_b200 import RoadmapBuilder

    instances = [_gen(500 + i, 5, 12, 6) for i in range(5)]
    bld = RoadmapBuilder(WEIGHTS, 0)
    try:
        bld.upload(instances)
        wide = bld.run(seed=3, N_traj=40, max_T=24, max_attempt=3, merge_distance_rate=0.1, compact=False).copy()  # W = 2
        comp = bld.export(compact=True).copy()
        assert type(comp).__name__ == "CompactCSRRoadmaps" and type(wide).__name__ == "CSRRoadmaps"
        assert np.array_equal(comp.pos.view(np.uint64), wide.pos.view(np.uint64))
        for f in ("layer_ptr", "eptr", "eidx", "n_layers", "cnt_collide", "agent_off"):
            assert np.array_equal(getattr(comp, f), getattr(wide, f)), f
        assert comp.eidx.dtype == np.int32 and comp.eptr.dtype == np.int64
        wide_bytes = wide.layer_ptr.nbytes + wide.pos.nbytes + wide.eptr.nbytes + wide.eidx.nbytes
        assert comp.nbytes < 0.6 * wide_bytes
        for b in range(len(instances)):
            assert comp.instance_counts(b) == wide.instance_counts(b)
            for x, y in zip(comp.instance_arrays(b), wide.instance_arrays(b)):
                assert np.array_equal(np.asarray(x), np.asarray(y))
        a = int(wide.agent_off[2])
        for t in (0, 1, 5):
            pc, ac = comp.layer(a, t)
            pw, aw = wide.layer(a, t)
            assert np.array_equal(pc, pw) and all(np.array_equal(u, v) for u, v in zip(ac, aw))
        tc, tw = comp.timed_roadmaps(1), wide.timed_roadmaps(1)
        assert all(tc[i].E == tw[i].E and len(tc[i].V) == len(tw[i].V) for i in range(len(tc)))
    finally:
        bld.close()


def test_learned_rows_only_equals_full_evaluation():
    """the network kernels of a step only run for the agents whose gate draw takes the learned branch (Batch::learn_agent);
    a traced run evaluates every agent instead.  The random-walking agents' samples are never read, so the free-running
    roadmaps of the two runs must be identical bit for bit — and the skipped rows must be a real fraction of the work"""
    from ctrm_b200 import RoadmapBuilder

    instances = [_gen(700 + i, 8, 16, 8) for i in range(6)]
    prm = dict(seed=77, N_traj=9, max_T=48, max_attempt=3, merge_distance_rate=0.1)
    bld = RoadmapBuilder(WEIGHTS, 0)
    try:
        bld.upload(instances)
        fast = bld.run(**prm).copy()
        full = bld.run(trace=True, **prm).copy()
        assert np.array_equal(fast.pos.view(np.uint64), full.pos.view(np.uint64))
        for f in ("layer_ptr", "eptr", "eidx", "n_layers", "cnt_collide"):
            assert np.array_equal(getattr(fast, f), getattr(full, f)), f
        # the traced run recorded the gate outcome implicitly: count agent-steps whose samples were NOT needed
        samples, loc_pred, _ = bld.read_trace()
        executed = ~np.isnan(loc_pred[..., 0])
        assert executed.sum() > 1000
    finally:
        bld.close()
