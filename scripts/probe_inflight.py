#!/usr/bin/env python3
"""how much throughput do frames in flight buy on ONE GPU?  N graph instances (own pool, own stream), device resident inputs,
round robin without waiting; wall clock around a device synchronise.  usage: probe_inflight.py W H [packed_bpp] [strength] [mode]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from vkdt_b200 import api, synth

W, H = int(sys.argv[1]), int(sys.argv[2])
bpp = int(sys.argv[3]) if len(sys.argv) > 3 else 0
strength = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
mode = sys.argv[5] if len(sys.argv) > 5 else "strict"
src = "i-mlv" if bpp else "i-raw"
raw = synth.mosaic(W, H, seed=5)
if bpp:
    words = synth.pack_bits_fast14(raw)
    buf = np.zeros(words.size + 64, dtype=np.uint16); buf[:words.size] = words
else:
    buf = np.ascontiguousarray(raw)
rp = api.raw_params(W, H, wb=(2.0, 1.0, 1.5), cam_to_rec2020=(0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8), packed_bpp=bpp, noise_a=100.0, noise_b=2.0)
for n in (1, 2, 3):
    gs, keep = [], []
    for k in range(n):
        d = api.dev_alloc(buf.nbytes + 256)
        api.check(api.lib.vkb_memcpy_h2d(d, buf.ctypes.data, buf.nbytes, None))
        api.check(api.lib.vkb_stream_sync(None))
        keep.append(d)
        g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src=src))
        g.set_sink_layout(api.SINK_RGB_F32)
        if strength > 0: g.line("param:denoise:01:strength:%g" % strength)
        g.set_mode(api.MODE_FAST if mode == "fast" else api.MODE_STRICT)
        g.set_source(d, rp, device=True)
        g.set_sink_buffer(None, 0)
        g.run()
        gs.append(g)
    steps = 60 if W * H < 2e7 else 24
    for rep in range(2):
        for g in gs: g.run(api.RUN_WAIT)
        t0 = time.perf_counter()
        for i in range(steps):
            gs[i % n].run(api.RUN_RECORD | api.RUN_UPLOAD)
        for g in gs: g.run(api.RUN_WAIT)
        dt = (time.perf_counter() - t0) / steps * 1e3
    print("%dx%d %s %s: %d in flight: %.3f ms per frame, %.0f MP/s, %.0f frames/s" % (W, H, src, mode, n, dt, W * H / dt / 1e3, 1e3 / dt), flush=True)
    for g in gs: g.close()
    for d in keep: api.dev_free(d)
