"""Stand-alone timing of the HBM-style stage kernels through the C-ABI (the kernels north_star asks to see against the HBM
roofline: occupancy raster, cost-to-go wavefront, FOV raster, swept-circle collision with survivor compaction).
Each kernel runs alone on working sets far larger than the 126 MB L2, timed with CUDA events on the launching stream
after warm-up; achieved = ALGORITHMIC bytes / time, peak = MEASURED_PEAKS.json.

    python tools/stage_bench.py [--md profiles/r1_stage_kernels.md]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WEIGHTS, make_instances  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--md", default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--quick", action="store_true", help="one un-warmed launch per kernel on smaller inputs (the target of an ncu capture)")
    args = ap.parse_args()
    import torch

    from ctrm_b200 import RoadmapBuilder, _lib as L

    lib = L.load()
    torch.cuda.set_device(0)
    st = torch.cuda.current_stream().cuda_stream
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    rows = []

    def timed(fn, iters=args.iters):
        if args.quick:
            iters = 1
        for _ in range(0 if args.quick else 2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def report(name, ms, nbytes, unit_desc, extra=""):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append((name, unit_desc, ms, nbytes / 1e6, gbs, gbs / hbm, extra))
        print(f"{name:28s} {ms:9.3f} ms  {nbytes / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {100 * gbs / hbm:5.1f}% of {hbm:.0f}  {extra}")

    M, F = 160, 19
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731

    # ---------------- instances of the benchmark shape
    B = 1024
    instances = make_instances(B, 0)
    p = RoadmapBuilder.pack_instances(instances)
    n_agents, n_obs = p["n_agents"].astype(np.int64), p["n_obs"].astype(np.int64)
    A = int(n_agents.sum())
    obs_off = dev(np.concatenate([[0], np.cumsum(n_obs)]).astype(np.int32))
    obs = dev(p["obs"])
    agent_inst = dev(np.repeat(np.arange(B), n_agents).astype(np.int32))

    # ---------------- occupancy raster: 8 x the batch so that the output (210 MB) exceeds L2
    rep = 8
    obs_r = dev(np.tile(p["obs"], (rep, 1)))
    off_r = dev(np.concatenate([[0], np.cumsum(np.tile(n_obs, rep))]).astype(np.int32))
    occ_r = torch.empty((rep * B, M, M), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: L.check(lib.ctrm_raster_occupancy(obs_r.data_ptr(), off_r.data_ptr(), rep * B, M, occ_r.data_ptr(), st)))
    report("raster_occupancy", ms, rep * B * M * M, f"{rep * B} maps of {M}x{M} u8 written")
    occ = occ_r[:B].contiguous()

    # ---------------- cost-to-go wavefront: one field per goal (51.2 KB written each)
    goals = dev(p["goals"])
    sr = dev(np.array([int(np.ceil(float(r) * M)) for r in p["rads"]], dtype=np.int32))
    ctg = torch.empty((A, M, M), dtype=torch.uint16, device="cuda")
    ms = timed(lambda: L.check(lib.ctrm_cost_to_go(occ.data_ptr(), agent_inst.data_ptr(), goals.data_ptr(), sr.data_ptr(), A, M,
                                                   ctg.data_ptr(), st)), iters=3)
    c16 = ctg.view(torch.int16)  # hop counts < 32768; 65535 (unreachable) reads as -1
    levels = int(c16.max().item())
    report("cost_to_go (wavefront)", ms, A * M * M * 2 + B * M * M, f"{A} goals, {M}x{M} u16 fields written",
           f"deepest field {levels} levels; host builds footprint stencils inside the call")

    # ---------------- FOV raster, stage API (f32 rows, row-major cost field): 8 positions per agent
    rep = 8
    pos = dev(np.tile(p["starts"], (rep, 1)) + 0.0)
    inst_r = agent_inst.repeat(rep)
    ctg_idx = None
    fov = torch.empty((rep * A, 2 * F * F), dtype=torch.float32, device="cuda")
    # the stage API addresses the cost field per agent row: replicate the index, not the 1.3 GB of fields
    ctg_r = ctg  # agent a of replica k reads field a: pass replica-major positions with a matching field table
    ms = 0.0
    for k in range(rep):  # one launch per replica (each reads all A fields once, writes A rows)
        pk = pos[k * A:(k + 1) * A]
        fk = fov[k * A:(k + 1) * A]
        ms += timed(lambda: L.check(lib.ctrm_fov_raster(occ.data_ptr(), ctg_r.data_ptr(), 0, agent_inst.data_ptr(), pk.data_ptr(), A, M,
                                                        F, fk.data_ptr(), st)), iters=3)
    ms /= rep
    report("fov_raster (f32 stage API)", ms, A * (2 * F * F * 4 + F * F * 3 + 16), f"{A} crops of 2x{F}x{F} f32 written",
           "3.97 KB algorithmic per crop (SURVEY 8d)")

    # ---------------- swept-circle collision with survivor compaction, O = 10 and O = 64 obstacles per instance
    rng = np.random.RandomState(0)
    for O_per in (10, 64):
        n = 1 << (22 if args.quick else 25)
        Bc = 4096
        ob = np.concatenate([rng.rand(Bc * O_per, 2), rng.uniform(0.025, 0.04, (Bc * O_per, 1))], axis=1)
        ooff = dev((np.arange(Bc + 1) * O_per).astype(np.int32))
        obd = dev(ob)
        f = torch.rand((n, 2), dtype=torch.float64, device="cuda")
        t = f + (torch.rand((n, 2), dtype=torch.float64, device="cuda") - 0.5) * (2.0 / 32.0)
        r = torch.full((n,), 1.0 / 64.0, dtype=torch.float64, device="cuda")
        si = torch.randint(0, Bc, (n,), dtype=torch.int32, device="cuda").sort().values.contiguous()
        verdict = torch.empty(n, dtype=torch.uint8, device="cuda")
        surv = torch.empty(n, dtype=torch.int32, device="cuda")
        ns = torch.zeros(1, dtype=torch.int32, device="cuda")
        ms = timed(lambda: L.check(lib.ctrm_collide_segments(f.data_ptr(), t.data_ptr(), r.data_ptr(), si.data_ptr(), n, obd.data_ptr(),
                                                             ooff.data_ptr(), verdict.data_ptr(), surv.data_ptr(), ns.data_ptr(), st)))
        frac_hit = 1.0 - ns.item() / n
        # fp64 work actually done: every (segment, obstacle) pair pays the bounding-box reject (2 adds for rad_sum + margin, 4
        # adds for the grown box = 6 flop, compares not counted); only pairs that pass it run the exact closest-approach test
        # (~40 flop).  Count the passing pairs with the same box test in torch, chunked.
        passed = 0
        ob3 = obd.reshape(Bc, O_per, 3)
        for c0 in range(0, n, 1 << 21):
            sl = slice(c0, min(n, c0 + (1 << 21)))
            o = ob3[si[sl].long()]                                  # (m, O, 3)
            R = (r[sl, None] + o[..., 2]) + 1e-9
            lox, hix = torch.minimum(f[sl, 0], t[sl, 0])[:, None], torch.maximum(f[sl, 0], t[sl, 0])[:, None]
            loy, hiy = torch.minimum(f[sl, 1], t[sl, 1])[:, None], torch.maximum(f[sl, 1], t[sl, 1])[:, None]
            far = (o[..., 0] < lox - R) | (o[..., 0] > hix + R) | (o[..., 1] < loy - R) | (o[..., 1] > hiy + R)
            passed += int((~far).sum().item())
        flops = n * O_per * 6.0 + passed * 40.0
        report(f"collide_segments O={O_per}", ms, n * (40 + 4 + 1 + 4), f"{n} swept checks x {O_per} obstacles, compaction",
               f"{flops / (ms * 1e-3) / 1e12:.2f} TFLOP/s fp64 actually executed ({passed / (n * O_per) * 100:.1f}% of the pairs pass the "
               f"box reject and run the 40-flop test, the rest cost 6 flop), {100 * frac_hit:.1f}% collide")
        del f, t, r, si, verdict, surv

    if args.md:
        with open(args.md, "w") as fh:
            fh.write("Stage kernels timed alone through the C-ABI (`tools/stage_bench.py`, CUDA events, working sets > L2); achieved = "
                     f"algorithmic bytes / time against the measured HBM peak of {hbm:.0f} GB/s (MEASURED_PEAKS.json)\n\n")
            fh.write("| kernel | work | ms | algorithmic MB | GB/s | frac of HBM peak | notes |\n|---|---|---:|---:|---:|---:|---|\n")
            for name, desc, ms, mb, gbs, fr, extra in rows:
                fh.write(f"| {name} | {desc} | {ms:.3f} | {mb:.1f} | {gbs:.0f} | {fr:.3f} | {extra} |\n")


if __name__ == "__main__":
    main()
