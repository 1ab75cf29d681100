"""print the per-launch timings (-d perf) of one warm run of the default graph at a given size."""
import sys
import numpy as np
sys.path.insert(0, ".")
from vkdt_b200 import api, synth
W, H = int(sys.argv[1]), int(sys.argv[2])
strength = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
api.init(0)
raw = synth.mosaic(W, H, seed=1, wb=(2.0, 1.0, 1.5))
g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
if strength > 0:
    g.line("param:denoise:01:strength:%g" % strength)
d = api.dev_alloc(raw.nbytes + 256)
api.check(api.lib.vkb_memcpy_h2d(d, raw.ctypes.data, raw.nbytes, None)); api.check(api.lib.vkb_stream_sync(None))
g.set_source(d, api.raw_params(W, H, wb=(2.0, 1.0, 1.5), noise_a=100.0, noise_b=2.0), device=True)
g.set_sink_layout(api.SINK_RGB_F32)  # as bench.py: the PFM payload
g.set_sink_buffer(None, 0)
g.run()
for _ in range(3):
    g.run(api.RUN_RECORD | api.RUN_UPLOAD | api.RUN_WAIT | api.RUN_PERF)
tot = 0.0
for label, ms, nb in g.perf_entries():
    tot += ms
    print("%-58s %8.3f ms %7.1f GB/s" % (label[:58], ms, nb / ms / 1e6 if ms > 0 else 0))
print("sum of launches %.3f ms" % tot)
