"""one line per kernel launch from the text summary scripts/ncu_summary.py writes (units normalised)."""
import sys
U = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
rows, cur = [], None
for l in open(sys.argv[1]).read().split("\n"):
    if not l.startswith("   "):
        if l.strip():
            cur = {"name": l.strip()}
            rows.append(cur)
    else:
        p = l.split()
        try:
            v = float(p[1].replace(",", ""))
        except ValueError:
            continue
        if len(p) > 2 and p[2] in U:
            v *= U[p[2]]
        cur[p[0]] = v
g = lambda r, k: r.get(k, float("nan"))
print("%-34s %8s %7s %7s %5s %5s %5s %4s %6s %5s %5s %5s %5s %9s | stalls: long short barr math mio lg" % (
    "kernel", "us", "rdMB", "wrMB", "sm%", "dram%", "occ%", "regs", "issue%", "fma%", "alu%", "xu%", "lsu%", "Minst"))
for r in rows:
    if g(r, "gpu__time_duration.sum") < float(sys.argv[2]) if len(sys.argv) > 2 else 0:
        continue
    s = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
    print("%-34s %8.1f %7.1f %7.1f %5.1f %5.1f %5.1f %4.0f %6.1f %5.1f %5.1f %5.1f %5.1f %9.1f | %5.2f %5.2f %5.2f %5.2f %5.2f %5.2f" % (
        r["name"][:34], g(r, "gpu__time_duration.sum"), g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum"),
        g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(r, "launch__registers_per_thread"),
        g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        g(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), g(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        g(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), g(r, "smsp__inst_executed.sum") / 1e6,
        g(r, s % "long_scoreboard"), g(r, s % "short_scoreboard"), g(r, s % "barrier"), g(r, s % "math_pipe_throttle"), g(r, s % "mio_throttle"), g(r, s % "lg_throttle")))
