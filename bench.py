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


def make_instances(n, seed0):
    """synthetic instances with the shape of the aamas22 'homo-basis' benchmark, seed 46+i"""
    import numpy as np

    from ctrm_b200.environment import generate_ins_2d_with_obs_hetero_nonfix_agents as gen

    return [gen(21, 30, "0.03125", "0.015625", 10, 0.05, 0.08, rng=np.random.RandomState(46 + seed0 + i)) for i in range(n)]


# ------------------------------------------------------------------------------------------------ CPU (oracle port)
def _cpu_one(args):
    seed, n_traj = args
    import numpy as np
    import torch

    torch.set_num_threads(1)
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


def cpu_worker(nproc, n_traj, count, seed0):
    """times the oracle (the CPU restatement of the reference's path: numpy + torch-CPU, 1 torch thread per process,
    one process per core — the best CPU throughput setting of BASELINE.md §3) on `count` instances"""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    with ctx.Pool(nproc) as pool:
        pool.map(_cpu_one, [(10 ** 6 + i, 1) for i in range(nproc)])  # warm the workers (imports, weights)
        t0 = time.perf_counter()
        res = pool.map(_cpu_one, [(seed0 + i, n_traj) for i in range(count)], chunksize=1)
        dt = time.perf_counter() - t0
    return dict(vertices=int(sum(r[0] for r in res)), edges=int(sum(r[1] for r in res)), seconds=dt, instances=count,
                nproc=nproc, n_traj=n_traj)


def run_cpu_subprocess(nproc, n_traj, count, seed0):
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker", "--nproc", str(nproc), "--cpu-ntraj",
                          str(n_traj), "--cpu-count", str(count), "--cpu-seed", str(seed0)], capture_output=True, text=True,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1"))
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
    core; each step is a bounded sample sized so that steps + warmup finish within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(host_cores(), 64)
    total = max(1, args.steps + args.warmup)
    budget = 150.0 / total                      # seconds per step
    n_traj = int(max(1, min(PARAMS["N_traj"], budget / 1.6)))   # ~1-1.5 s per roll-out per core, growing with the roadmap
    vals, last = [], None
    for i in range(total):
        r = run_cpu_subprocess(cores, n_traj, cores, 5000 + 1000 * i)
        if i >= args.warmup:
            vals.append(r)
        last = r
    v = sum(r["vertices"] for r in vals) / sum(r["seconds"] for r in vals)
    ms = 1e3 * sum(r["seconds"] for r in vals) / len(vals)
    sample = f"{cores} instances of the workload per step, one per core, N_traj={n_traj} of 25 roll-outs (bounded sample)"
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64+f32", data="synthetic",
                config=dict(workload=WORKLOAD, instances_per_step=cores, n_traj_sample=n_traj),
                instances_per_s=sum(r["instances"] for r in vals) / sum(r["seconds"] for r in vals) * n_traj / PARAMS["N_traj"],
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
def algorithmic_work(kernel, agent_steps, pair_steps, K, F=19):
    """ALGORITHMIC bytes / flops of one kernel over the counted units (DESIGN.md 'Kernels')"""
    FF = F * F
    if kernel == "fov_raster":      # write 2F^2 bf16 {0,1} operand-tile entries, read F^2 (1 B occ + 2 B cost) + position
        return dict(bound="hbm", work=agent_steps * (2 * FF * 2 + FF * 3 + 16), unit="GB/s")
    if kernel == "fov_encode":      # 2 x (722->32->32->32) + the 32x32 vain projection
        return dict(bound="tensor", work=agent_steps * 2.0 * (2 * FF * 64 + 4 * 1024 + 1024), unit="TFLOP/s")
    if kernel == "agent_comm":      # per ordered pair: 11x32 + 32x32 + 32x42 MAC
        return dict(bound="tensor", work=pair_steps * 2.0 * (11 * 32 + 1024 + 32 * 42), unit="TFLOP/s")
    if kernel == "cvae_prior":      # indicator 3,424 + encoder_input 5,472 + decoder x-part 75x32 MAC per agent
        return dict(bound="tensor", work=agent_steps * 2.0 * (3424 + 5472 + 75 * 32), unit="TFLOP/s")
    if kernel == "cvae_decode":     # z-part of layer 0 + layers 1, 2 per (agent, copy) row
        return dict(bound="tensor", work=agent_steps * K * 2.0 * (64 * 32 + 1024 + 96), unit="TFLOP/s")
    # step (select + merge + advance): per agent-step K samples x 12 B + state ~96 B + <= 3 layers x 26 candidates x
    # (16 B position + 8 B masks) + the new vertex
    return dict(bound="hbm", work=agent_steps * (K * 12 + 96 + 3 * 26 * 24 + 32), unit="GB/s")


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this workload at 1,024
# instances (profiles/r1_full_step_kernels_b1024_final.md); scaled linearly with the instance count
NCU_TRAFFIC_B1024 = dict(fov_raster=59.7e6, fov_encode=20.2e6, agent_comm=8.42e6, cvae_prior=7.69e6, cvae_decode=10.1e6,
                         step=14.7e6)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # FMA issue peak of the CUDA cores at the max SM clock


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
        out = gather_to_rank0(builder.export_device())
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
    nv, ne, _ = builder.result_sizes()
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

    # ---- end to end through the host API (host packing + H2D of the instance arrays and D2H of the CSR inside the timed
    #      region, every step): a stream of batches through PipelinedBuilder, which overlaps one batch's copies with the
    #      other's kernels (two contexts on this GPU, alternating)
    from ctrm_b200 import PipelinedBuilder

    e2e_steps = max(2, args.steps)
    pipe = PipelinedBuilder(WEIGHTS, local)
    for _ in pipe.map([instances] * 2, seed=seed, instance_id_base=rank * B, **PARAMS):
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
    d2h = int(csr.layer_ptr.nbytes + csr.pos.nbytes + csr.eptr.nbytes + csr.eidx.nbytes + csr.n_layers.nbytes + csr.cnt_collide.nbytes)
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
    for name, v in kt.items():
        if not v["launches"]:
            continue
        w = algorithmic_work(name, agent_steps, pair_steps, 3)
        achieved = w["work"] / (v["ms"] * 1e-3) / (1e9 if w["bound"] == "hbm" else 1e12)
        peak = hbm_peak if w["bound"] == "hbm" else tc_peak
        kernels[name] = dict(ms_per_launch=v["ms"] / v["launches"], share=v["ms"] / total_kernel_ms, bound=w["bound"],
                             achieved=achieved, peak=peak, unit=w["unit"], frac=achieved / peak,
                             traffic=NCU_TRAFFIC_B1024.get(name, 0.0) * B / 1024.0 or None)
        if w["bound"] == "tensor":
            kernels[name]["frac_of_fp32_fma_peak"] = achieved / FP32_PEAK_TFLOPS
    dom = max(kernels, key=lambda k: kernels[k]["share"])
    roofline = dict(kernel=dom, bound=kernels[dom]["bound"], achieved=kernels[dom]["achieved"], peak=kernels[dom]["peak"],
                    unit=kernels[dom]["unit"], frac=kernels[dom]["frac"], traffic=kernels[dom]["traffic"],
                    share_of_step=kernels[dom]["share"], peak_source=peak_src,
                    frac_of_fp32_fma_peak=kernels[dom].get("frac_of_fp32_fma_peak"),
                    note="achieved = algorithmic work of all launches / summed CUDA-event time of the kernel (eager pass over "
                         "the same workload right after the timed region, one CUDA event after every kernel on the driver's "
                         "stream); traffic = ncu dram bytes per launch at 1,024 instances scaled to this batch; the MLP kernels run "
                         "fp32-exact BF16x3 / fixed-point digit products on tcgen05 (6-10 MMAs per useful one) and are bound by "
                         "their epilogues and load latency, not by the tensor pipe; they are reported against the dense bf16 "
                         "tensor peak as the contract asks, with the fp32 FMA peak beside it")

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
                config=dict(workload=WORKLOAD, instances_per_gpu=B, instances_per_step=int(b_all),
                            l2="working set (cost-to-go fields + roadmap store, ~3.5 MB/instance) exceeds the 126 MB L2",
                            parallelism=f"instances sharded over {world} GPU(s), final NCCL gather" if world > 1 else "1 GPU"),
                instances_per_s=b_all / (ms_per_step * 1e-3), ms_per_instance=ms_per_step / b_all * world,
                vertices_per_step=nv_all, edges_per_step=ne_all,
                device_ms=dict(total=tim["total_ms"], rollout=tim["rollout_ms"], finalize=tim["finalize_ms"], graph_replays=tim["steps"]),
                clocks=clocks, e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                                        ms_per_step=e2e_t[0].item(), steps=e2e_steps,
                                        api="PipelinedBuilder.map (2 contexts, copies of one batch overlap the other's kernels)"),
                gpu_launches=int(launches_all), roofline=roofline, kernels=kernels, cpu_baseline=cpu, planner=planner)
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
    ap.add_argument("--batch", type=int, default=int(os.environ.get("CTRM_BENCH_BATCH", "2960")), help="instances per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-worker", action="store_true")
    ap.add_argument("--nproc", type=int, default=1)
    ap.add_argument("--cpu-ntraj", type=int, default=25)
    ap.add_argument("--cpu-count", type=int, default=1)
    ap.add_argument("--cpu-seed", type=int, default=0)
    args = ap.parse_args()
    if args.cpu_worker:
        print(json.dumps(cpu_worker(args.nproc, args.cpu_ntraj, args.cpu_count, args.cpu_seed)))
        return
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
