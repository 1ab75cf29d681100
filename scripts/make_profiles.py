"""assemble profiles/rNN_* from the ncu outputs a gpurun call left in gpurun_out/ (see profiles/README.md for the commands).
usage: python scripts/make_profiles.py r01 gpurun_out/r01_launches.csv gpurun_out/bench.json summary1.txt [summary2.txt ...]"""
import collections
import csv
import json
import shutil
import subprocess
import sys

tag, launches_csv, bench_json = sys.argv[1:4]
summaries = sys.argv[4:]
out = "profiles/"
shutil.copy(launches_csv, out + tag + "_launches.csv")
with open(out + tag + "_ncu_full_summary.txt", "w") as f:
    for s in summaries:
        f.write(open(s).read())

# ---- share of one step per kernel, from the serialised ncu launch list ----
rows = list(csv.reader(open(launches_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, rows = rows[hi], rows[hi + 1:]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
names = [r[ki].split("(")[0].replace("void ", "") for r in rows]
vals = [float(r[vi]) / 1e3 for r in rows]  # us
first = [i for i, n in enumerate(names) if n == names[0]]
period = first[1] - first[0]
step = first[len(first) // 2]               # a warm step in the middle of the run
ncu_share = collections.OrderedDict()
for n, v in zip(names[step:step + period], vals[step:step + period]):
    ncu_share[n] = ncu_share.get(n, 0.0) + v
ncu_total = sum(ncu_share.values())

bench = json.load(open(bench_json))
ev = bench["roofline"]["kernels"]
LABEL = {"k_denoise_half": "denoise_half", "k_denoise_downcov": "denoise_downcov", "k_denoise_down_tiled": "denoise_down", "k_denoise_down": "denoise_down",
         "k_denoise_assemble": "denoise_assemble", "k_denoise_doub_bayer": "denoise_doub", "k_hilite_half": "hilite_half", "k_hilite_reduce": "hilite_reduce",
         "k_hilite_assemble": "hilite_assemble", "k_hilite_doub": "hilite_doub", "k_demosaic_gauss<0>": "demosaic_gauss", "k_bayer_splat": "demosaic_splat",
         "k_bayer_fix": "demosaic_fix", "k_llap_reduce0<1>": "b200_llapr0", "k_llap_reduce0_p<1>": "b200_llapr0", "k_llap_reduce": "llap_reduce", "k_llap_assemble_tiled": "llap_assemble",
         "k_llap_assemble": "llap_assemble", "k_llap_assemble4": "llap_assemble", "k_llap_final4<1, 1>": "b200_llapfin", "k_pointwise_t<1, 2, 3, 0, 0, 1>": "b200_pointw"}

# ---- per launch table of the full capture ----
tab = subprocess.run([sys.executable, "scripts/ncu_table.py", out + tag + "_ncu_full_summary.txt", "0"], capture_output=True, text=True).stdout

# ---- dram traffic per launch of the biggest launch of each kernel ----
U = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic, issue, cur = {}, {}, None
for l in open(out + tag + "_ncu_full_summary.txt"):
    if not l.startswith("   "):
        cur = l.strip().split("(")[0].replace("void ", "")
        continue
    p = l.split()
    if p[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        traffic.setdefault(cur, []).append(float(p[1].replace(",", "")) * U.get(p[2], 1.0))
    if p[0] == "smsp__issue_active.avg.pct_of_peak_sustained_active":
        issue.setdefault(cur, []).append(float(p[1].replace(",", "")))
tr, iss = {}, {}
for k, v in traffic.items():
    pairs = [v[i] + v[i + 1] for i in range(0, len(v) - 1, 2)]
    big = max(range(len(pairs)), key=lambda i: pairs[i])
    tr[LABEL.get(k, k)] = pairs[big]
    if k in issue and big < len(issue[k]):
        iss[LABEL.get(k, k)] = issue[k][big]
json.dump({"comment": "largest launch of each kernel, ncu --set full, 9504x6336 still, denoise 0.4: dram__bytes_read.sum + dram__bytes_write.sum, "
                      "and smsp__issue_active.avg.pct_of_peak_sustained_active",
           "bytes_per_launch": tr, "sm_issue_pct": iss}, open(out + tag + "_traffic.json", "w"), indent=1)

with open(out + tag + "_summary.md", "w") as f:
    f.write("# %s: ncu summary of the default darkroom graph, 9504x6336 bayer still, denoise strength 0.4\n\n" % tag)
    f.write("bench line of the same build (not under the profiler): value %.1f %s, %.3f ms/step, e2e %.1f %s.\n\n" % (
        bench["value"], bench["unit"], bench["ms_per_step"], bench["e2e"]["value"], bench["e2e"]["unit"]))
    f.write("## share of one step per kernel: ncu launch list (serialised, cold cache) vs CUDA events inside bench.py\n\n")
    f.write("| kernel | launches | ncu us | ncu share | events: largest launch ms |\n|---|---|---|---|---|\n")
    cnt = collections.Counter(names[step:step + period])
    for n, v in sorted(ncu_share.items(), key=lambda kv: -kv[1]):
        lab = LABEL.get(n, n)
        e = [x["ms"] for k, x in ev.items() if lab in k.split()]
        f.write("| %s | %d | %.1f | %.3f | %s |\n" % (n, cnt[n], v, v / ncu_total, ("%.3f" % e[0]) if e else "-"))
    f.write("\nncu total of the step %.3f ms (sum of serialised launches); bench ms_per_step %.3f; dominant kernel share per bench %.3f.\n" % (
        ncu_total / 1e3, bench["ms_per_step"], bench["roofline"]["share_of_step"]))
    f.write("\n## ncu --set full, one line per launch (largest launches of each kernel first in pyramid order)\n\n```\n" + tab + "```\n")
print(open(out + tag + "_summary.md").read()[:3000])
