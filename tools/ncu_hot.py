"""Per source line: executed warp instructions and stall samples of one kernel, from an .ncu-rep captured with
--import-source on (the library is built with -lineinfo).  usage: python tools/ncu_hot.py <rep> [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
inst, samp, text = collections.Counter(), collections.Counter(), {}
cur = None
H = None
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        H = r
        ii, si = H.index("Instructions Executed"), H.index("# Samples")
        continue
    if H is None or len(r) < len(H):
        continue
    if r[0].strip():  # a source line row, followed by its SASS rows
        cur = int(r[0])
        text[cur] = r[1].strip()
        continue
    if cur is None:
        continue
    try:
        inst[cur] += int(r[ii] or 0)
        samp[cur] += int(r[si] or 0)
    except ValueError:
        pass
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print(f"total warp instructions {ti}, samples {ts}")
for line, n in inst.most_common(top):
    print(f"  line {line:4d}  inst {100 * n / ti:5.1f}%  stall samples {100 * samp[line] / ts:5.1f}%  {text.get(line, '')[:110]}")
