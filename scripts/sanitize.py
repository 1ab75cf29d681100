"""run a few graph variants at awkward sizes: meant to be started under compute-sanitizer (memcheck / racecheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from vkdt_b200 import api, synth
api.init(0)
for (W, H, xt, strength, method, layout) in [(530, 412, False, 0.4, 0, 1), (261, 195, True, 0.4, 0, 0), (70, 50, False, 0.4, 0, 1),
                                              (646, 412, False, 0.0, 1, 0), (516, 408, True, 0.0, 2, 1), (2050, 1030, False, 0.4, 0, 1)]:
    raw = np.ascontiguousarray(synth.mosaic(W, H, seed=3, xtrans=xt))
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    if strength > 0:
        g.line("param:denoise:01:strength:%g" % strength)
    g.line("param:demosaic:01:method:%d" % method)
    g.set_source(raw.ctypes.data, api.raw_params(W, H, wb=(2.0, 1.0, 1.5), noise_a=100.0, noise_b=2.0, **({"filters": 9} if xt else {})))
    g.set_sink_layout(layout)
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    out = np.zeros((oh, ow, 3 if layout else 4), np.float32)
    g.set_sink_buffer(out.ctypes.data, out.nbytes)
    g.run(api.RUN_RECORD | api.RUN_UPLOAD | api.RUN_DOWNLOAD | api.RUN_WAIT)
    print(W, H, "xtrans" if xt else "bayer", strength, method, layout, "->", ow, oh, float(out.mean()), bool(np.isfinite(out).all()))
    g.close()
# packed mlv frame
W, H = 1024, 514
raw = synth.mosaic(W, H, seed=5)
words = synth.pack_bits_fast14(raw) if raw.size % 8 == 0 else synth.pack_bits(raw, 14)
buf = np.zeros(words.size + 64, dtype=np.uint16); buf[:words.size] = words
g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-mlv"))
g.set_source(buf.ctypes.data, api.raw_params(W, H, wb=(2.0, 1.0, 1.5), packed_bpp=14))
g.set_sink_buffer(None, 0)
g.run()
print("mlv ok", g.sink_size())
g.close()
# round 2: sized export (resize: catmull-rom, blur + slice, flower), 8 bit export, dng gain maps (noop and doub), both builds
import struct
for (mw, mh, sink, strength, gm, mode) in [(300, 0, "o-pfm", 0.4, True, api.MODE_STRICT), (100, 100, "o-jpg", 0.0, True, api.MODE_STRICT),
                                           (1200, 1200, "o-pfm", 0.0, False, api.MODE_FAST), (0, 0, "o-jpg", 0.4, False, api.MODE_FAST)]:
    W, H = 642, 484
    raw = np.ascontiguousarray(synth.mosaic(W, H, seed=9))
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"), sink=sink, prim=1 if sink == "o-jpg" else None, trc=1 if sink == "o-jpg" else None,
                  max_width=mw, max_height=mh)
    g.set_mode(mode)
    if strength > 0:
        g.line("param:denoise:01:strength:%g" % strength)
    g.set_source(raw.ctypes.data, api.raw_params(W, H, wb=(2.0, 1.0, 1.5), noise_a=100.0, noise_b=2.0))
    if gm:
        ops = [synth.dng_gain_map_opcode(np.full((5, 7), 1.0 + 0.1 * k, np.float32), k >> 1, k & 1, H, W) for k in range(4)]
        g.set_dng_opcodes(synth.dng_opcode_list(ops))
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    out = np.zeros((oh, ow, 4), np.uint8 if sink == "o-jpg" else np.float32)
    g.set_sink_buffer(out.ctypes.data, out.nbytes)
    g.run(api.RUN_RECORD | api.RUN_UPLOAD | api.RUN_DOWNLOAD | api.RUN_WAIT)
    print("export", mw, mh, sink, strength, gm, mode, "->", ow, oh, float(out.mean()))
    g.close()
