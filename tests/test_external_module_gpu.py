"""a module that is not part of the library (include/vkdt_b200.h, "module authors"): connectors + params as text, one kernel,
registered through the C-ABI, named in a cfg like any other module, run inside the default darkroom graph between colour and
filmcurv.  the kernel is compiled here with nvcc from tests/tools/ext_module.cu."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from helpers import to_dev_f16, dev_f16, to_host, fbits
from vkdt_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    so = str(tmp_path / "libinvert.so")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "tools", "ext_module.cu")], check=True)
    return C.CDLL(so)


def test_external_module_builds_and_registers_on_the_host(tmp_path):
    """host half (no GPU): the module's files register, a cfg names it, the planner schedules its kernel as a launch."""
    from vkdt_b200 import api
    ext = _build(tmp_path)
    conn = C.c_char_p.in_dll(ext, "invert_connectors").value
    par = C.c_char_p.in_dll(ext, "invert_params").value
    api.lib.vkb_register_module.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    api.lib.vkb_register_kernel.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int]
    api.check(api.lib.vkb_register_module(b"invert", conn, par))
    api.check(api.lib.vkb_register_kernel(b"invert", b"main", C.cast(ext.invert_main, C.c_void_p), -1))
    assert "param amount:float:1:0:0000803f" in api.module_describe("invert")
    cfg = api.DARKROOM_CFG.format(src="i-raw").replace("connect:colour:01:output:filmcurv:01:input\n",
        "module:invert:01\nconnect:colour:01:output:invert:01:input\nconnect:invert:01:output:filmcurv:01:input\nparam:invert:01:amount:0.75\n")
    g = api.Graph(cfg_text=cfg)
    raw = np.zeros((384, 512), dtype=np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(512, 384))
    plan = g.plan()
    assert "invert_main" in plan and "b200_pointw (crop+colour)" in plan, plan
    g.close()


@pytest.mark.gpu
def test_external_module_runs_inside_the_graph(gpu, oracle, tmp_path):
    ext = _build(tmp_path)
    gpu.lib.vkb_register_module.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    gpu.lib.vkb_register_kernel.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int]
    gpu.check(gpu.lib.vkb_register_module(b"invert", C.c_char_p.in_dll(ext, "invert_connectors").value, C.c_char_p.in_dll(ext, "invert_params").value))
    gpu.check(gpu.lib.vkb_register_kernel(b"invert", b"main", C.cast(ext.invert_main, C.c_void_p), -1))
    w, h = 512, 384
    raw = synth.mosaic(w, h, seed=3)
    WB, CAM = (2.0, 1.0, 1.5), (0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8)
    # the oracle up to colour (stage 5), the module's arithmetic in numpy, the rest of the oracle's graph by hand
    d = oracle.darkroom_defaults(w, h)
    for k in range(3): d.whitebalance[k] = WB[k]
    for k in range(9): d.cam_to_rec2020[k] = CAM[k]
    col = oracle.darkroom_run(d, raw, stage=5)
    inv = col.copy()
    inv[..., :3] = (np.float32(0.75) - col[..., :3]).astype(np.float16).astype(np.float32)
    O = oracle
    flm, fi = O.new_img(inv.shape[0], inv.shape[1], 4)
    O.lib().o_filmcurv_main(C.byref(O.img(inv)), C.byref(fi), C.byref(d.filmcurv), 1)
    ll, li = O.new_img(inv.shape[0], inv.shape[1], 4)
    O.lib().o_llap_module(C.byref(O.img(flm)), C.byref(li), C.byref(d.llap), 1)
    want, wi = O.new_img(inv.shape[0], inv.shape[1], 4)
    O.lib().o_grade_main(C.byref(O.img(ll)), C.byref(wi), C.byref(d.grade), 0)
    cfg = gpu.DARKROOM_CFG.format(src="i-raw").replace("connect:colour:01:output:filmcurv:01:input\n",
        "module:invert:01\nconnect:colour:01:output:invert:01:input\nconnect:invert:01:output:filmcurv:01:input\nparam:invert:01:amount:0.75\n")
    g = gpu.Graph(cfg_text=cfg)
    buf = np.ascontiguousarray(raw)
    g.set_source(buf.ctypes.data, gpu.raw_params(w, h, wb=WB, cam_to_rec2020=CAM))
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    out = np.zeros((oh, ow, 4), dtype=np.float32)
    g.set_sink_buffer(out.ctypes.data, out.nbytes)
    g.set_perf(True)
    g.run(gpu.RUN_RECORD | gpu.RUN_UPLOAD | gpu.RUN_DOWNLOAD | gpu.RUN_WAIT)
    assert "invert_main" in g.perf()
    assert out.shape == want.shape
    assert np.array_equal(out[..., :3], want[..., :3]), float(np.abs(out[..., :3] - want[..., :3]).max())
    g.close()
