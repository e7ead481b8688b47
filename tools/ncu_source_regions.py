"""Per-kernel view of an ncu --set full --import-source capture: stall samples and executed instructions between
synchronisation points (barriers, mbarrier waits, TMEM loads, MMA issue), to see which phase of a kernel holds the time.
    python tools/ncu_source_regions.py gpurun_out/prof.ncu-rep [kernel-regex]
"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
MARK = re.compile(r"BAR\.|SYNCS\.PHASECHK|LDTM|UTCBAR|UTCHMMA|EXIT|UBLKCP|RET\.")
for blk in blocks:
    if pat and not pat.search(blk["name"]):
        continue
    hdr, data = blk["rows"][0], blk["rows"][1:]
    i_s, i_ex, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot_s = sum(int(r[i_s] or 0) for r in data) or 1
    tot_e = sum(int(r[i_ex] or 0) for r in data) or 1
    print(f"\n## {blk['name'][:60]}  samples {tot_s}  instructions {tot_e / 1e6:.2f}M")
    cs = ce = 0
    for i, r in enumerate(data):
        cs += int(r[i_s] or 0)
        ce += int(r[i_ex] or 0)
        if MARK.search(r[i_src]) and (cs or ce):
            print(f"  ..{i:5d}  samples {cs:5d} ({100 * cs / tot_s:4.1f}%)  inst {ce / 1e6:6.2f}M ({100 * ce / tot_e:4.1f}%)  {r[i_src].strip()[:60]}")
            cs = ce = 0
    if cs or ce:
        print(f"  ..end    samples {cs:5d} ({100 * cs / tot_s:4.1f}%)  inst {ce / 1e6:6.2f}M")
