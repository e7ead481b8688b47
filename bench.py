#!/usr/bin/env python
"""bench.py — timed-roadmap construction throughput on B200 (BASELINE.json metric) with the reference CPU path beside it.

  python bench.py --gpus N --steps K --warmup W            one rank per GPU (torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU path (oracle port) on the host cores

A "step" is one pass of the hot path over one batch of synthetic aamas22-shape instances (configs[1] of BASELINE.json:
21-30 agents, 10 circular obstacles, trained aamas22-main networks, K = 3 candidates, N_traj = 25, max_T = 64, eval
parameters of scripts/conf/eval/eval_large_ctrm_learned_ind.yaml): cost-to-go for every goal, all roll-outs, finalisation
and the CSR export.  `value` = vertices/s with the instance definitions already resident on the device and the result left
in HBM; `e2e` = the same metric through the public host API (host instance arrays in, host CSR out, copies inside).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
WEIGHTS = os.path.join(ROOT, "tests", "golden", "aamas22_main_best.npz")
METRIC = "timed_roadmap_vertices_per_s"
UNIT = "vertices/s"
PARAMS = dict(N_traj=25, max_T=64, max_attempt=3, merge_distance_rate=0.1, prob_uniform_sampling_after_goal=0.9,
              prob_uniform_bias=0.0, prob_uniform_gamma=5.0)
WORKLOAD = "aamas22 homo-basis: 21-30 agents, 10 obstacles, rad 1/64, speed 1/32, K=3, N_traj=25, max_T=64, merge_rate 0.1"
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")  # written by tools/summarize_ncu.py traffic
STRONG_TOTAL = 1024  # BASELINE configs[3]: 1,024 instances of the workload sharded over the GPUs of the box


def config_dict(batch, world):
    """the `config` of the JSON line: both arms print the same dict (the reference arm runs a bounded sample of it)"""
    return dict(workload=WORKLOAD, instances_per_gpu=int(batch), instances_per_step=int(batch) * int(world),
                l2="working set (cost-to-go fields + roadmap store, ~3.5 MB/instance) exceeds the 126 MB L2",
                parallelism=f"instances sharded over {world} GPU(s), final NCCL gather" if world > 1 else "1 GPU")


def make_instances(n, seed0, n_lo=21, n_hi=30, obs=10, rads="0.015625", lo=0.05, hi=0.08):
    """synthetic instances with the shape of the aamas22 'homo-basis' benchmark, seed 46+i;
    n_lo/n_hi = 31/40 gives 'homo-more-agents' (config 3), 100/101 with 64 small obstacles the stress shape (config 5)"""
    import numpy as np

    from ctrm_b200.environment import generate_ins_2d_with_obs_hetero_nonfix_agents as gen

    return [gen(n_lo, n_hi, "0.03125", rads, obs, lo, hi, rng=np.random.RandomState(46 + seed0 + i)) for i in range(n)]


# ------------------------------------------------------------------------------------------------ CPU (oracle port)
def _cpu_one(args):
    seed, n_traj = args[:2]
    import numpy as np
    import torch

    if len(args) < 3 or args[2] > 0:  # 0 = leave torch's default intra-op threads (what the notebooks did)
        torch.set_num_threads(args[2] if len(args) > 2 else 1)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ctrm_oracle as O
    from rng_tables import PhiloxTables

    ins = make_instances(1, seed)[0]
    s, g, sp, r, o = ins.arrays()
    oi = O.OInstance.from_arrays(s, g, sp, r, o[:, :2], o[:, 2])
    w = O.OWeights.load(WEIGHTS)
    p = O.OParams(N_traj=n_traj, max_T=PARAMS["max_T"], max_attempt=PARAMS["max_attempt"],
                  merge_distance_rate=PARAMS["merge_distance_rate"])
    trms = O.build_timed_roadmaps(oi, w, p, PhiloxTables(1234, seed))
    nv, ne = O.trm_counts(trms)
    return nv, ne


def cpu_worker(nproc, n_traj, count, seed0, threads=1, steps=1, warm=1):
    """times the oracle (the CPU restatement of the reference's path: numpy + torch-CPU, 1 torch thread per process,
    one process per core — the best CPU throughput setting of BASELINE.md §3) on `count` instances"""
    import multiprocessing as mp

    if nproc == 1:  # the single-process settings of BASELINE.md §3: 1 torch thread, or torch's default thread count (threads=0)
        _cpu_one((10 ** 6, 1, threads))
        t0 = time.perf_counter()
        res = [_cpu_one((seed0 + i, n_traj, threads)) for i in range(count)]
        dt = time.perf_counter() - t0
        return dict(vertices=int(sum(r[0] for r in res)), edges=int(sum(r[1] for r in res)), seconds=dt, instances=count,
                    nproc=1, n_traj=n_traj, threads=threads)

    ctx = mp.get_context("fork")
    out = []
    with ctx.Pool(nproc) as pool:
        for w in range(max(1, warm)):  # warm the workers (imports, weights, allocator): one short roll-out each, untimed
            pool.map(_cpu_one, [(10 ** 6 + 1000 * w + i, 1) for i in range(nproc)])
        for k in range(steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_one, [(seed0 + 1000 * k + i, n_traj) for i in range(count)], chunksize=1)
            dt = time.perf_counter() - t0
            out.append(dict(vertices=int(sum(r[0] for r in res)), edges=int(sum(r[1] for r in res)), seconds=dt, instances=count,
                            nproc=nproc, n_traj=n_traj))
    return out[0] if steps == 1 else dict(steps=out)


def run_cpu_subprocess(nproc, n_traj, count, seed0, threads=1, steps=1, warm=1):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    if threads > 0:
        env.update(OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads))
    else:
        env.pop("OMP_NUM_THREADS", None)
        env.pop("MKL_NUM_THREADS", None)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker", "--nproc", str(nproc), "--cpu-ntraj",
                          str(n_traj), "--cpu-count", str(count), "--cpu-seed", str(seed0), "--cpu-threads", str(threads),
                          "--cpu-steps", str(steps), "--cpu-warm", str(warm)],
                         capture_output=True, text=True, env=env)
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("cpu worker failed: " + out.stderr[-2000:])


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  /root/reference (Python + numba + two pybind11
    helpers) cannot travel to the GPU box, so this times the oracle port of it (kind "port") with one process per host
    core, every process building ONE instance of the workload at the FULL N_traj = 25 (13-20 s per step): the headline
    config, a bounded number of instances."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(host_cores(), 128)
    n_traj = PARAMS["N_traj"]
    # one pool for the whole run; the W warm-up steps are short roll-outs (they only have to warm imports, weights and the
    # allocator of every worker and are not timed), the K timed steps run the full workload
    r = run_cpu_subprocess(cores, n_traj, cores, 5000, steps=max(1, args.steps), warm=max(1, args.warmup))
    vals = r["steps"] if "steps" in r else [r]
    v = sum(r["vertices"] for r in vals) / sum(r["seconds"] for r in vals)
    ms = 1e3 * sum(r["seconds"] for r in vals) / len(vals)
    sample = (f"{cores} instances of the workload per step at the full N_traj={n_traj}, one per host core (one process per core, "
              f"1 torch thread each); warm-up steps are single roll-outs")
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64+f32", data="synthetic",
                config=config_dict(args.batch, args.gpus),
                instances_per_s=sum(r["instances"] for r in vals) / sum(r["seconds"] for r in vals),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled every 200 ms DURING the timed region, in-process through NVML (a polling
    nvidia-smi subprocess was seen to perturb kernel launches); falls back to `nvidia-smi -lms` when NVML is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.sm, self.mx, self.reasons, self.proc, self.stop_flag, self.mode = gpu_index, [], [], set(), None, False, None

    def _nvml_loop(self):
        import pynvml as N

        h = N.nvmlDeviceGetHandleByIndex(self._nvml_index)
        bits = {"hw_slowdown": getattr(N, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(N, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        self.mx.append(float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)))
        while not self.stop_flag:
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        try:
            import pynvml as N

            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            self._nvml_index = int(vis.split(",")[self.gpu]) if vis and vis.replace(",", "").isdigit() else self.gpu
            self.mode = "nvml"
            self.th = threading.Thread(target=self._nvml_loop, daemon=True)
            self.th.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 9 and r[1].replace(".", "").isdigit():
                self.sm.append(float(r[1]))
                if r[2].replace(".", "").isdigit():
                    self.mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)

    def stop(self):
        self.stop_flag = True
        if self.mode is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock sampling unavailable"])
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        else:
            self.th.join(timeout=1.0)
        sm = sorted(self.sm)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(self.mx) if self.mx else None,
                    reasons=sorted(self.reasons), samples=len(sm), source=self.mode)


# ------------------------------------------------------------------------------------------------ GPU arm
def algorithmic_work(kernel, agent_steps, pair_steps, K, F=19, avg_layer=20.0):
    """ALGORITHMIC bytes / flops of one kernel over the counted units (DESIGN.md 'Kernels')"""
    FF = F * F
    if kernel == "fov_raster":      # write 2F^2 BITS (96 B per row), read F^2 (1 B occ + 2 B cost) + position
        return dict(bound="hbm", work=agent_steps * (96 * ((2 * FF + 767) // 768) + FF * 3 + 16), unit="GB/s")
    if kernel == "fov_encode":      # 2 x (722->32->32->32) + the 32x32 vain projection
        return dict(bound="tensor", work=agent_steps * 2.0 * (2 * FF * 64 + 4 * 1024 + 1024), unit="TFLOP/s")
    if kernel == "agent_comm":      # per ordered pair: 11x32 + 32x32 + 32x42 MAC
        return dict(bound="tensor", work=pair_steps * 2.0 * (11 * 32 + 1024 + 32 * 42), unit="TFLOP/s")
    if kernel == "cvae_prior":      # indicator 3,424 + encoder_input 5,472 + decoder x-part 75x32 MAC per agent
        return dict(bound="tensor", work=agent_steps * 2.0 * (3424 + 5472 + 75 * 32), unit="TFLOP/s")
    if kernel == "cvae_decode":     # z-part of layer 0 + layers 1, 2 per (agent, copy) row
        return dict(bound="tensor", work=agent_steps * K * 2.0 * (64 * 32 + 1024 + 96), unit="TFLOP/s")
    if kernel == "cvae":            # the fused prior + decode kernel
        return dict(bound="tensor", work=agent_steps * 2.0 * (3424 + 5472 + 75 * 32) + agent_steps * K * 2.0 * (64 * 32 + 1024 + 96),
                    unit="TFLOP/s")
    # step (select + merge + advance) per agent-step: K samples x 12 B, agent / instance state ~112 B, the positions of the
    # three layers the merge looks at (V[t-1], V[t], V[t+1]: on average half of their final size while the roadmap grows,
    # 16 B each), the masks of the few merge candidates and the written vertex / masks / next position ~72 B
    return dict(bound="hbm", work=agent_steps * (K * 12 + 112 + 3 * 0.5 * avg_layer * 16 + 72), unit="GB/s")


def ncu_traffic(batch):
    """dram bytes per launch per kernel from the tracked ncu capture (profiles/r2_ncu_traffic.json, tools/summarize_ncu.py
    traffic), scaled linearly from the capture's instance count to this batch; None when the file is absent"""
    try:
        doc = json.load(open(TRAFFIC_JSON))
        return {k: v["dram_bytes_per_launch"] * batch / float(doc["instances"]) for k, v in doc["kernels"].items()}, doc["source"]
    except Exception:
        return {}, None


FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # FMA issue peak of the CUDA cores at the max SM clock


def other_configs(builder, device, seed):
    """BASELINE configs[0], [2] and [4] on one GPU, device-resident timing (CUDA events of ctrm_build_roadmaps + upload wall
    time), each after one warm-up pass"""
    import numpy as np

    from ctrm_b200 import RoadmapBuilder
    from ctrm_b200.environment import Instance

    out = {}

    def timed(bld, instances, reps, **prm):
        packed = bld.pack_instances(instances)
        best = None
        for r in range(reps + 1):
            t0 = time.perf_counter()
            bld.upload(None, packed=packed)
            csr = bld.run(seed=seed, **prm)
            wall = (time.perf_counter() - t0) * 1e3
            if r > 0 and (best is None or wall < best[0]):
                best = (wall, bld.timing(), csr.n_vertices, csr.n_edges)
        wall, tim, nv, ne = best
        return dict(instances=len(instances), agents=int(packed["n_agents"].sum()), ms_wall=wall, ms_device=tim["total_ms"],
                    graph_replays=tim["steps"], vertices=nv, edges=ne, vertices_per_s=nv / (wall * 1e-3),
                    instances_per_s=len(instances) / (wall * 1e-3))

    # configs[0]: the demo instance (notebooks/CTRM_demo_ins.pkl as arrays), ONE instance per call like the reference's
    # roadmap_gen (scripts/eval.py:57-61): host arrays in, host CSR out — the single-instance latency
    z = np.load(os.path.join(ROOT, "tests", "golden", "demo_ins.npz"))
    demo = Instance.from_arrays(z["starts"], z["goals"], z["max_speeds"], z["rads"], z["obs_pos"], z["obs_rad"])
    for tag, rate in (("demo_instance", 0.25), ("demo_instance_eval_params", 0.1)):
        r = timed(builder, [demo], 3, **dict(PARAMS, merge_distance_rate=rate))
        r.update(latency_ms=r["ms_wall"], merge_distance_rate=rate,
                 note="28 agents, 10 obstacles, N_traj=25, max_T=64; wall clock of upload + build + export for ONE instance")
        out[tag] = r
    # configs[2]: homo-more-agents, 31-40 agents
    more = make_instances(1480, 200000, n_lo=31, n_hi=40)
    out["more_agents_31_40"] = timed(builder, more, 1, **PARAMS)
    del more
    # configs[4]: the stress shape — 100 agents, 64 obstacles, K = 64 candidates, map_size 250, N_traj = 25, max_T = 64
    stress = make_instances(64, 300000, n_lo=100, n_hi=101, obs=64, rads="0.00390625", lo=0.02, hi=0.04)
    sb = RoadmapBuilder(WEIGHTS, device, map_size=250)
    try:
        out["stress_100a_64o_k64_m250"] = timed(sb, stress, 1, **dict(PARAMS, K=64))
    finally:
        sb.close()
    return out


def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from ctrm_b200 import RoadmapBuilder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the JSON result: anything a library prints there meanwhile (NCCL's version banner)
    # is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ctrm_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    instances = make_instances(B, rank * B)
    builder = RoadmapBuilder(WEIGHTS, local)
    packed = builder.pack_instances(instances)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    seed = 2022

    pending = []  # (works, keepalive) of the previous step's gather

    def drain_gather():
        while pending:
            works, keep = pending.pop()
            for w in works:
                w.wait()
            del keep

    def gather_to_rank0(res):
        """the only collective of the path: NCCL gather of the variable-length CSR payload over NVLink.  It is enqueued
        asynchronously: the transfer of step i overlaps the kernels of step i + 1 and is waited for before the next gather
        starts (and, for the last step, before the timed region closes)"""
        if world == 1:
            return res
        from ctrm_b200.parallel import gather_csr

        drain_gather()
        out, works, keep = gather_csr(res, rank, world, async_op=True)
        pending.append((works, (keep, res)))
        return out

    phase_ms = []

    def device_step():
        t0 = time.perf_counter()
        builder.upload(None, packed=packed)
        t1 = time.perf_counter()
        builder.run(seed=seed, instance_id_base=rank * B, export=False, **PARAMS)
        t2 = time.perf_counter()
        out = gather_to_rank0(builder.export_device(compact=world > 1))
        phase_ms.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (time.perf_counter() - t2)))
        return out

    def e2e_step():
        csr = builder.build(instances, seed=seed, instance_id_base=rank * B, **PARAMS)
        return csr

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        device_step()
    drain_gather()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record()
    for _ in range(args.steps):
        device_step()
        launches += builder.timing()["launches"] + 5  # + occupancy (2), passability, wavefront, csr fill
    drain_gather()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if rank == 0:
        sys.stderr.write("host ms per step (upload, build, export_device+gather): " + ", ".join(
            "(%.0f %.0f %.0f)" % p for p in phase_ms[-args.steps:]) + "\n")
    clocks = sampler.stop() if rank == 0 else None
    nv, ne, host_n_layers_sum = builder.result_sizes()
    steps_inst = builder.instance_steps().astype(np.int64)
    n_ag = packed["n_agents"].astype(np.int64)
    tim = builder.timing()
    stats = torch.tensor([ms, float(nv), float(ne), float(B), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, nv_all, ne_all, b_all, launches_all = mx[0].item(), sm[1].item(), sm[2].item(), sm[3].item(), sm[4].item()
    else:
        nv_all, ne_all, b_all, launches_all = float(nv), float(ne), float(B), float(launches)
    ms_per_step = ms / args.steps
    value = nv_all / (ms_per_step * 1e-3)

    # ---- strong scaling (BASELINE configs[3]): STRONG_TOTAL instances of the workload sharded over the ranks, same timing
    #      discipline (barrier + sync on both sides, max over ranks), final gather inside
    from ctrm_b200.parallel import shard_range

    s_lo, s_hi = shard_range(STRONG_TOTAL, rank, world)
    s_inst = instances[:s_hi - s_lo] if world == 1 and B >= STRONG_TOTAL else make_instances(s_hi - s_lo, 100000 + s_lo)
    s_packed = builder.pack_instances(s_inst)

    def strong_step():
        builder.upload(None, packed=s_packed)
        builder.run(seed=seed, instance_id_base=100000 + s_lo, export=False, **PARAMS)
        return gather_to_rank0(builder.export_device(compact=world > 1))

    for _ in range(2):
        strong_step()
    drain_gather()
    barrier()
    strong_steps = max(3, min(args.steps, 10))
    ev0.record()
    for _ in range(strong_steps):
        strong_step()
    drain_gather()
    ev1.record()
    barrier()
    s_nv, _, _ = builder.result_sizes()
    s_stats = torch.tensor([ev0.elapsed_time(ev1) / strong_steps, float(s_nv)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = s_stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = s_stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        s_stats = torch.stack([mx[0], sm[1]])
    strong = dict(scaling="strong", instances_total=STRONG_TOTAL, instances_per_gpu=s_hi - s_lo, n_gpus=world, steps=strong_steps,
                  ms_per_step=s_stats[0].item(), value=s_stats[1].item() / (s_stats[0].item() * 1e-3), unit=UNIT,
                  instances_per_s=STRONG_TOTAL / (s_stats[0].item() * 1e-3),
                  note="BASELINE configs[3]: 1,024 aamas22-shape instances in total, sharded with shard_range, gather inside; "
                       "device-resident like `value`")
    del s_inst, s_packed

    # ---- end to end through the host API (host packing + H2D of the instance arrays and D2H of the CSR inside the timed
    #      region, every step): a stream of batches through PipelinedBuilder, which overlaps one batch's copies with the
    #      other's kernels (two contexts on this GPU, alternating)
    from ctrm_b200 import PipelinedBuilder

    e2e_steps = max(6, args.steps)  # a stream of batches: the fill and drain of the pipeline are 1 / e2e_steps of it
    depth = 3 if world == 1 else 2  # per rank; every context owns pinned staging for one batch's result (4.4 GB)
    pipe = PipelinedBuilder(WEIGHTS, local, depth=depth)
    for _ in pipe.map([instances] * depth, seed=seed, instance_id_base=rank * B, **PARAMS):
        pass
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    e2e_nv = 0
    for csr in pipe.map([instances] * e2e_steps, seed=seed, instance_id_base=rank * B, **PARAMS):
        e2e_nv += csr.n_vertices
    ev1.record()
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    h2d = int(sum(v.nbytes for v in packed.values()))
    d2h = int(csr.nbytes)  # CompactCSRRoadmaps: bytes of the arrays ctrm_export_csr_compact copied to the host
    pipe.close()
    e2e_t = torch.tensor([e2e_ms, -e2e_ms, float(e2e_nv) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = e2e_t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = e2e_t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_t = torch.stack([mx[0], mx[1], sm[2]])
    e2e_value = e2e_t[2].item() / (e2e_t[0].item() * 1e-3)

    # ---- per-kernel device times (eager launches + CUDA events on the driver's stream), same workload
    builder.set_profile(True)
    builder.upload(None, packed=packed)
    builder.run(seed=seed, instance_id_base=rank * B, export=False, **PARAMS)
    kt = builder.kernel_times()
    builder.set_profile(False)
    agent_steps = float((steps_inst * n_ag).sum())
    pair_steps = float((steps_inst * n_ag * (np.minimum(15, n_ag - 1) + 1)).sum())  # self + k nearest per agent
    total_kernel_ms = sum(v["ms"] for v in kt.values()) or 1.0
    kernels = {}
    traffic, traffic_src = ncu_traffic(B)
    avg_layer = float(nv) / max(1.0, float(host_n_layers_sum))
    for name, v in kt.items():
        if not v["launches"]:
            continue
        w = algorithmic_work(name, agent_steps, pair_steps, 3, avg_layer=avg_layer)
        achieved = w["work"] / (v["ms"] * 1e-3) / (1e9 if w["bound"] == "hbm" else 1e12)
        peak = hbm_peak if w["bound"] == "hbm" else tc_peak
        kernels[name] = dict(ms_per_launch=v["ms"] / v["launches"], share=v["ms"] / total_kernel_ms, bound=w["bound"],
                             achieved=achieved, peak=peak, unit=w["unit"], frac=achieved / peak,
                             traffic=traffic.get(name))
        if w["bound"] == "tensor":
            kernels[name]["frac_of_fp32_fma_peak"] = achieved / FP32_PEAK_TFLOPS
    dom = max(kernels, key=lambda k: kernels[k]["share"])
    roofline = dict(kernel=dom, bound=kernels[dom]["bound"], achieved=kernels[dom]["achieved"], peak=kernels[dom]["peak"],
                    unit=kernels[dom]["unit"], frac=kernels[dom]["frac"], traffic=kernels[dom]["traffic"],
                    share_of_step=kernels[dom]["share"], peak_source=peak_src, traffic_source=traffic_src,
                    frac_of_fp32_fma_peak=kernels[dom].get("frac_of_fp32_fma_peak"),
                    note="achieved = algorithmic work of all launches / summed CUDA-event time of the kernel (eager pass over "
                         "the same workload right after the timed region, one CUDA event after every kernel on the driver's "
                         "stream); traffic = ncu dram bytes per launch read from profiles/r2_ncu_traffic.json (tools/summarize_ncu.py traffic), "
                         "scaled linearly from the capture's instance count to this batch; the MLP kernels run "
                         "fp32-exact BF16x3 / fixed-point digit products on tcgen05 (6-10 MMAs per useful one) and are bound by "
                         "their epilogues and load latency, not by the tensor pipe; they are reported against the dense bf16 "
                         "tensor peak as the contract asks, with the fp32 FMA peak beside it")

    # ---- the other BASELINE configs on one GPU (extra keys; not part of `value`)
    other = None
    if rank == 0 and world == 1 and not args.no_extra:
        other = other_configs(builder, local, seed)

    # ---- the roadmaps' consumer: prioritized planning of the whole batch on the device-resident CSR
    csr_dev = builder.export_device()
    builder.plan(csr_dev)
    pr = builder.plan(csr_dev)
    planner = dict(ms_per_batch=pr["elapsed_ms"], instances=B, instances_per_s=B / (pr["elapsed_ms"] * 1e-3),
                   solved_frac=float(pr["solved"].float().mean().item()),
                   collision_checks=int(pr["counters"][:, 0].sum().item()), explored=int(pr["counters"][:, 2].sum().item()),
                   note="ctrm_plan_prioritized on this rank's last batch, CUDA events; not part of `value`")
    pp_host = None
    if rank == 0 and world == 1 and not args.no_cpu:  # one SOLVED instance of that batch, kept for the planner's CPU leg
        from ctrm_b200.planner import _flatten_trms

        sol = pr["solved"].cpu().numpy()
        pp_b = int(np.nonzero(sol)[0][0]) if sol.any() else 0
        host_csr = builder.export()
        pp_host = [np.array(x) for x in _flatten_trms(host_csr.timed_roadmaps(pp_b))[1:]]
        pp_L = host_csr.L
        planner["cpu_port_instance_explored"] = int(pr["counters"][pp_b, 2].item())
        del host_csr
    del csr_dev, pr
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = min(host_cores(), 64)
        r = run_cpu_subprocess(cores, PARAMS["N_traj"], cores, 9000)
        cpu = dict(value=r["vertices"] / r["seconds"], unit=UNIT, cores=cores, kind="port",
                   sample=f"{cores} instances of the same workload at full N_traj=25, one process per core, oracle port "
                          f"(numpy + torch-CPU, 1 thread each), {r['seconds']:.1f} s wall",
                   instances_per_s=r["instances"] / r["seconds"],
                   single_core_s_per_instance=r["seconds"] * cores / r["instances"])  # one instance per process: its latency
        # BASELINE.md section 3's other two settings, on ONE instance with a bounded N_traj = 8 of 25 roll-outs: one process
        # with 1 torch thread (best single-core latency) and one process with torch's default threads (what the notebooks did)
        for tag, th in (("one_process_1_thread", 1), ("one_process_default_threads", 0)):
            r1 = run_cpu_subprocess(1, 8, 1, 9500, threads=th)
            cpu[tag] = dict(vertices_per_s=r1["vertices"] / r1["seconds"], seconds=r1["seconds"],
                            sample="1 instance, N_traj=8 of 25 (bounded sample)")
        # planner, CPU port (oracle/pp_oracle.py), one instance of the same workload on one core
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pp_oracle

        n_layers, lp, pos, eptr, eidx = pp_host
        lp_flat = np.concatenate([lp[a * pp_L:a * pp_L + int(n)] for a, n in enumerate(n_layers)] + [lp[-1:]])
        ins0 = instances[pp_b].arrays()
        t0 = time.perf_counter()
        got = pp_oracle.prioritized_planning(ins0[0], ins0[1], ins0[2], ins0[3], pp_oracle.FlatRoadmaps(n_layers, lp_flat, pos, eptr, eidx))
        planner["cpu_port_s_per_instance"] = time.perf_counter() - t0
        planner["cpu_port_solved"] = bool(got["solved"])
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64+f32",
                data="synthetic",
                config=config_dict(B, world),
                instances_per_s=b_all / (ms_per_step * 1e-3), ms_per_instance=ms_per_step / b_all * world,
                vertices_per_step=nv_all, edges_per_step=ne_all,
                device_ms=dict(total=tim["total_ms"], rollout=tim["rollout_ms"], finalize=tim["finalize_ms"], graph_replays=tim["steps"]),
                clocks=clocks, e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                                        ms_per_step=e2e_t[0].item(), steps=e2e_steps,
                                        api=f"PipelinedBuilder.map ({depth} contexts on the GPU, the first with a high-priority stream: packing, "
                                            "copies and export of one batch overlap the roll-out kernels of another)"),
                gpu_launches=int(launches_all), roofline=roofline, kernels=kernels, cpu_baseline=cpu, planner=planner,
                strong_scaling=strong, other_configs=other,
                latency_ms_single_instance=(other or {}).get("demo_instance", {}).get("latency_ms"))
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("CTRM_BENCH_BATCH", "5920")), help="instances per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra single-GPU configs (demo instance, 31-40 agents, stress)")
    ap.add_argument("--cpu-worker", action="store_true")
    ap.add_argument("--nproc", type=int, default=1)
    ap.add_argument("--cpu-ntraj", type=int, default=25)
    ap.add_argument("--cpu-count", type=int, default=1)
    ap.add_argument("--cpu-seed", type=int, default=0)
    ap.add_argument("--cpu-threads", type=int, default=1)
    ap.add_argument("--cpu-steps", type=int, default=1)
    ap.add_argument("--cpu-warm", type=int, default=1)
    args = ap.parse_args()
    if args.cpu_worker:
        print(json.dumps(cpu_worker(args.nproc, args.cpu_ntraj, args.cpu_count, args.cpu_seed, args.cpu_threads, args.cpu_steps,
                                    args.cpu_warm)))
        return
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
