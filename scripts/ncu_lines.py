"""executed warp instructions and stall samples per source line, per kernel, from an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_lines.py report.ncu-rep [kernel-substring] [top N]"""
import collections
import csv
import subprocess
import sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern, cur_file, cur_line = None, None, None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kern = r[1].split("(")[0].replace("void ", "")
        continue
    if r[0] == "Line No":
        continue
    if r[0] != "":
        cur_line = (kern, cur_file, int(r[0]))
        agg.setdefault(cur_line, [r[1], 0, 0])
        continue
    try:
        agg[cur_line][1] += int(r[7])
        agg[cur_line][2] += int(r[4])
    except (ValueError, IndexError, KeyError):
        pass
kernels = collections.OrderedDict()
for (k, f, l), v in agg.items():
    kernels.setdefault(k, []).append((f, l, v))
for k, lines in kernels.items():
    if want not in k:
        continue
    tot = sum(v[1] for _, _, v in lines) or 1
    ts = sum(v[2] for _, _, v in lines) or 1
    print("== %s: %.1f M warp instructions, %d samples" % (k, tot / 1e6, ts))
    for f, l, (src, n, smp) in sorted(lines, key=lambda t: -t[2][1])[:topn]:
        print("  %-20s %4d %6.2f%% inst %6.2f%% stall  %s" % (f, l, 100.0 * n / tot, 100.0 * smp / ts, src.strip()[:120]))
