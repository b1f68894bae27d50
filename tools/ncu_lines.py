"""Join an `ncu --page source --csv --print-source sass` dump with `nvdisasm -g` line info: stall samples per source line.
usage: python tools/ncu_lines.py <sass.csv> <cubin> <kernel-name-substring> [top]"""
import csv
import re
import subprocess
import sys
from collections import Counter

sass_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur, infn = None, False
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
samples, insts = Counter(), Counter()
for r in rows[2:]:
    if len(r) <= isamp:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    key = line_of.get(a - base)
    samples[key] += int(r[isamp] or 0)
    insts[key] += int(r[iinst] or 0)
tot, toti = sum(samples.values()), sum(insts.values())
print(f"total samples {tot}, warp instructions {toti}")
for key, s in samples.most_common(top):
    print(f"{100 * s / tot:5.1f}% samples  {100 * insts[key] / toti:5.1f}% instr  {key}")
