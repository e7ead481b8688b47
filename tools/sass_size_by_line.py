"""Code size of one kernel by source line (outermost inlining site), from nvdisasm -g of a cubin / object built with -lineinfo.
    python tools/sass_size_by_line.py ctrm_b200/build/instance.o 'instance_rollout_kernelILi1E' [bucket]
A per-instance kernel whose loop body does not fit the instruction cache re-fetches its code every roll-out step."""
import collections
import re
import subprocess
import sys
import tempfile
import os

obj, pat = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout
fn, cur, cnt = None, None, collections.Counter()
for l in txt.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    if "//## File" in l:
        sites = re.findall(r'"([^"]+)", line (\d+)', l)
        f, ln = sites[-1]
        cur = (f.split("/")[-1], int(ln) // bucket * bucket, sites[0][0].split("/")[-1] if len(sites) > 1 else "")
        continue
    if fn and pat in fn and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l):
        cnt[cur[:2] if cur else None] += 1
tot = sum(cnt.values())
print(f"{tot} instructions = {tot * 16 / 1024:.0f} KB")
for k, v in sorted(cnt.items(), key=lambda kv: (kv[0] is None, kv[0])):
    print(k, v)
