#!/usr/bin/env python
"""Hot SASS instructions of a kernel in an .ncu-rep: ``python tools/ncu_hot.py rep [top]``."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first row: kernel name, second: header
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > iexec and r[isamp].isdigit()]
tot_s = sum(int(r[isamp]) for r in data)
tot_e = sum(int(r[iexec]) for r in data)
print(f"{len(data)} instructions, {tot_s} samples, {tot_e} warp-instructions executed")
print("--- by samples")
for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
    print(f"{r[ia][-5:]} {int(r[isamp]):7d} {100 * int(r[isamp]) / tot_s:5.1f}%  exec {int(r[iexec]):10d}  {r[isrc][:90]}")
print("--- by executed count")
for r in sorted(data, key=lambda r: -int(r[iexec]))[:top // 2]:
    print(f"{r[ia][-5:]} exec {int(r[iexec]):10d} {100 * int(r[iexec]) / tot_e:5.1f}%  {r[isrc][:90]}")
