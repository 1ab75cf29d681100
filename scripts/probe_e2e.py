"""where does an end to end 61 MP frame spend its time?  phases of one graph instance timed with events on its stream (upload,
launches, download), then the two instance ping-pong of bench.py with per frame wall clock.  python scripts/probe_e2e.py"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from vkdt_b200 import api, synth

W, H = 9504, 6336
api.init(0)
raw = synth.mosaic(W, H, seed=1, wb=(2.0, 1.0, 1.5))
rp = api.raw_params(W, H, wb=(2.0, 1.0, 1.5), noise_a=100.0, noise_b=2.0)
hin = api.host_alloc(raw.nbytes + 64)
api.C.memmove(hin, raw.ctypes.data, raw.nbytes)


def graph():
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    g.line("param:denoise:01:strength:0.4")
    g.set_sink_layout(api.SINK_RGB_F32)
    g.set_source(hin, rp)
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    ho = api.host_alloc(ow * oh * 12)
    g.set_sink_buffer(ho, ow * oh * 12)
    return g, ho, ow * oh * 12


g, ho, nb = graph()
ev = [api.Event() for _ in range(4)]
for it in range(3):
    st = g.stream()
    ev[0].record(st); g.run(api.RUN_UPLOAD | api.RUN_RECORD * 0)
    ev[1].record(st); g.run(api.RUN_RECORD)
    ev[2].record(st); g.run(api.RUN_DOWNLOAD)
    ev[3].record(st); g.run(api.RUN_WAIT); ev[3].sync()
    print("sequential: upload %.2f ms, launches %.2f ms, download %.2f ms (%.1f GB/s)" % (
        ev[0].elapsed_ms(ev[1]), ev[1].elapsed_ms(ev[2]), ev[2].elapsed_ms(ev[3]), nb / ev[2].elapsed_ms(ev[3]) / 1e6))
g2, ho2, _ = graph()
gs = [g, g2]
FE = api.RUN_RECORD | api.RUN_UPLOAD | api.RUN_DOWNLOAD
for n in (12, 40):
    for k in range(2):
        gs[k].run(FE | api.RUN_WAIT)
    t0 = time.time(); stamps = []
    for i in range(n):
        gk = gs[i % 2]
        if i >= 2:
            gk.run(api.RUN_WAIT)
        stamps.append(time.time() - t0)
        gk.set_source(hin, rp)
        gk.run(FE)
    for k in range(2):
        gs[(n + k) % 2].run(api.RUN_WAIT)
    t = time.time() - t0
    d = np.diff(np.array(stamps)) * 1e3
    print("ping-pong %d frames: %.2f ms per frame; steady state issue intervals (ms): %s" % (n, t / n * 1e3, np.round(d[4:14], 2)))
# download only, back to back, from both instances: what the copy engine does without anything else on the GPU
for k in range(2):
    gs[k].run(api.RUN_WAIT)
t0 = time.time()
for i in range(10):
    gs[i % 2].run(api.RUN_DOWNLOAD)
for k in range(2):
    gs[k].run(api.RUN_WAIT)
t = time.time() - t0
print("downloads only, two streams: %.2f ms per frame (%.1f GB/s)" % (t / 10 * 1e3, nb * 10 / t / 1e9))
