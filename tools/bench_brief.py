"""Compact view of a bench.py JSON line (stdin): step time, throughput and the per-kernel launch times.
    python bench.py ... 2>&1 | python tools/bench_brief.py
"""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        if line.startswith("host ms"):
            print(line)
        continue
    d = json.loads(line)
    print(f"ms_per_step {d['ms_per_step']:.1f}  inst/s {d.get('instances_per_s', 0):.0f}  value {d['value']:.3e} {d['unit']}  "
          f"e2e {d.get('e2e', {}).get('value', 0):.3e}  device_ms {d.get('device_ms')}")
    for k, v in (d.get("kernels") or {}).items():
        extra = f" fp32frac {v['frac_of_fp32_fma_peak']:.3f}" if "frac_of_fp32_fma_peak" in v else ""
        print(f"  {k:12s} {1e3 * v['ms_per_launch']:7.1f} us  share {v['share']:.3f}  {v['achieved']:.1f}/{v['peak']:.0f} {v['unit']} "
              f"frac {v['frac']:.4f}{extra}")
    print("clocks", d.get("clocks"))
    e = d.get("e2e") or {}
    if e:
        print(f"e2e ms/step {e.get('ms_per_step', 0):.1f}  h2d {e.get('h2d_bytes_per_step', 0) / 1e6:.1f} MB  d2h {e.get('d2h_bytes_per_step', 0) / 1e6:.1f} MB")
    if d.get("strong_scaling"):
        s = d["strong_scaling"]
        print(f"strong {s['instances_total']} instances on {s['n_gpus']} GPU(s): {s['ms_per_step']:.1f} ms/step, {s['value']:.3e} {s['unit']}, "
              f"{s['instances_per_s']:.0f} inst/s")
    for k, v in (d.get("other_configs") or {}).items():
        print(f"  {k:28s} {v['instances']:5d} inst {v['agents']:6d} agents  wall {v['ms_wall']:8.1f} ms  device {v['ms_device']:8.1f} ms  "
              f"{v['vertices_per_s']:.3e} v/s  {v['instances_per_s']:.1f} inst/s")
    if d.get("cpu_baseline"):
        print("cpu_baseline", {k: v for k, v in d["cpu_baseline"].items() if k != "sample"})
    if d.get("planner"):
        print("planner", {k: v for k, v in d["planner"].items() if k != "note"})
