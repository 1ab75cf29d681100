"""ad-hoc: feed the oracle's stage k output to the CUDA stage k+1 and report where they part."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")  # run from the repo root: python tests/tools/debug_stages.py W H
from vkdt_b200 import api, synth
from oracle import oracle_py as O
import plans
from helpers import to_dev_f16, to_dev_u16, dev_f16, to_host, fbits, ibits, ubits, f16_ulp_diff

w, h = int(sys.argv[1]), int(sys.argv[2])
api.init(0)
raw = synth.mosaic(w, h, seed=11)
d = O.darkroom_defaults(w, h)
for k, v in enumerate((2.0, 1.0, 1.5)): d.whitebalance[k] = v
st = {s: O.darkroom_run(d, raw, s) for s in range(1, 8)}
fin = O.darkroom_run(d, raw)
def rep(name, got, want):
    e = np.abs(got - want)
    u = f16_ulp_diff(got, want)
    idx = np.unravel_index(np.argmax(e), e.shape)
    print("%-10s max abs %.3g at %s (got %.5g want %.5g) max ulp %d, >2ulp %.2e" % (name, e.max(), idx, got[idx], want[idx], u.max(), (u > 2).mean()))
# hilite on oracle denoise out
got = to_host(plans.hilite(api, to_dev_f16(st[1]), w, h, (0.985, 0.3, 0.6), wb=(2.0, 1.0, 1.5, 1.0)))
rep("hilite", got, st[2])
got, cov, green = plans.demosaic(api, to_dev_f16(st[2]), w, h)
rep("demosaic", to_host(got), st[3])
ow, oh = O.darkroom_out_size(d)
fc = np.zeros(20, dtype=np.float32)
O.lib().o_crop_commit(0, w, h, d.crop.perspect, d.crop.crop, C.byref(C.c_float(d.crop.rotate)), O.fptr(fc))
f = np.zeros(242, dtype=np.float32)
wb = (C.c_float * 4)(*d.colour.white)
O.lib().o_colour_commit(C.byref(d.colour), wb, d.whitebalance, d.cam_to_rec2020, d.colour_primaries, d.colour_trc, O.fptr(f))
d_out = dev_f16(oh, ow, 4)
api.dispatch("b200", "pointw", [api.image(to_dev_f16(st[3]), w, h, 4, "f16"), api.image(d_out, ow, oh, 4, "f16")], ubits(3, 1, 2, 3),
             fc.tobytes() + f.tobytes() + bytes(d.filmcurv))
rep("pointwise", to_host(d_out)[..., :3], st[6][..., :3])
gp = bytes(d.grade)
got = to_host(plans.llap(api, to_dev_f16(st[6]), ow, oh, (0.12, 1.0, 1.0, 0.2), grade=gp, out_f32=True))
rep("llap+grade", got[..., :3], fin[..., :3])
