"""Summaries of ncu outputs for profiles/ (run here, no GPU needed).
    python tools/summarize_ncu.py launches gpurun_out/launches.csv          > profiles/<name>.md
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep              > profiles/<name>.md
    python tools/summarize_ncu.py traffic gpurun_out/prof.ncu-rep <instances> > profiles/<name>.json
        per-kernel DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a --set full capture taken at
        <instances> instances per batch: the file bench.py reads `roofline.traffic` from, so the bench line and the ncu
        capture cannot disagree
"""
import json
import collections
import csv
import io
import re
import subprocess
import sys

FULL_METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
                "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
                "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
                "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
                "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_no_instructions",
                "smsp__pcsamp_warps_issue_stalled_selected"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"ncu launch list `{path}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"| {k} | {cnt[k]} | {v:.1f} | {v / cnt[k]:.2f} | {v / T:.3f} |")
    print(f"\ntotal {T:.0f} us over {sum(cnt.values())} launches")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"ncu --set full capture `{path}` (one launch per kernel)\n")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "")
        print(f"### {name}\n\n| metric | value | unit |\n|---|---:|---|")
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        print()


BENCH_KERNELS = (("fov_raster", r"fov_raster_tiles"), ("fov_encode", r"fov_encode_tc"), ("agent_comm", r"agent_comm_tc"),
                 ("cvae_prior", r"cvae_prior_tc"), ("cvae_decode", r"cvae_decode_tc"), ("cvae", r"cvae_tc_kernel"),
                 ("step", r"step(_thread)?_kernel"), ("advance", r"advance_kernel"), ("wavefront", r"wavefront_kernel"),
                 ("csr_fill", r"csr_fill_kernel"))


def traffic(path, instances):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "ns ": 1e-3}
    acc = collections.defaultdict(lambda: [0.0, 0.0, 0])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        kid = next((k for k, pat in BENCH_KERNELS if re.search(pat, name)), None)
        if kid is None:
            continue
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            b += float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        i = hdr.index("gpu__time_duration.sum")
        t = float(r[i].replace(",", "")) * tscale.get(units[i], 1.0)
        acc[kid][0] += b
        acc[kid][1] += t
        acc[kid][2] += 1
    doc = dict(source=path.split("/")[-1], instances=int(instances), note="ncu --set full --clock-control none, one replayed "
               "launch per kernel (cold cache, serialised); dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum per launch",
               kernels={k: dict(dram_bytes_per_launch=v[0] / v[2], ncu_us_per_launch=v[1] / v[2], launches=v[2]) for k, v in acc.items()})
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
