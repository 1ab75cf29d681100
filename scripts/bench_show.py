"""prints the headline numbers of a bench.py json line (gpurun_out/bench_<tag>.json)"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f %s  %.3f ms/step  e2e %.1f" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"]))
print("roofline:", d["roofline"]["kernel"], d["roofline"]["frac"])
print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
if "fast" in d: print("fast ms/step", d["fast"].get("ms_per_step"))
