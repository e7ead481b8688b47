"""Per-source-line stall samples / executed instructions of one kernel from an ncu --set full --import-source capture.
    python tools/ncu_source_lines.py gpurun_out/prof.ncu-rep kernel-regex [top]
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows, path, hdr = [], None, None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        num = lambda v: int(v) if v.strip().isdigit() else 0
        rows.append((path, int(r[0]), r[1].strip(), num(r[i_s]), num(r[i_e])))
ts, te = sum(r[3] for r in rows) or 1, sum(r[4] for r in rows) or 1
print(f"{kern}: {ts} samples, {te / 1e6:.2f}M instructions")
for p, ln, src, s, e in sorted(rows, key=lambda r: -r[3])[:top]:
    print(f"{100 * s / ts:5.1f}% smp {100 * e / te:5.1f}% inst  {p}:{ln:<4d} {src[:100]}")
