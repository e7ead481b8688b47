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
