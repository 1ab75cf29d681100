"""feedback connectors (connector.h:63 s_conn_feedback, double buffering graph-run-modules.h:169-196, frame crossing
graph-run-nodes-allocate.h:222-229): an outside module ("iir", tests/tools/ext_module.cu) reads its own output of the frame
before through a `feedback:` line.  host half: the traversal still reaches it, its output is double buffered, the input is
wired to the other copy.  gpu half: four frames of the recurrence inside the darkroom graph against numpy + the oracle."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from vkdt_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG_EDIT = ("connect:colour:01:output:filmcurv:01:input\n",
            "module:iir:01\nconnect:colour:01:output:iir:01:input\nfeedback:iir:01:output:iir:01:back\n"
            "connect:iir:01:output:filmcurv:01:input\nparam:iir:01:keep:0.25\n")


def _register(api, tmp_path):
    so = str(tmp_path / "libext.so")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "tools", "ext_module.cu")], check=True)
    ext = C.CDLL(so)
    api.lib.vkb_register_module.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    api.lib.vkb_register_kernel.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int]
    api.check(api.lib.vkb_register_module(b"iir", C.c_char_p.in_dll(ext, "iir_connectors").value, C.c_char_p.in_dll(ext, "iir_params").value))
    api.check(api.lib.vkb_register_kernel(b"iir", b"main", C.cast(ext.iir_main, C.c_void_p), -1))
    return ext


def test_feedback_plan_on_the_host(tmp_path):
    from vkdt_b200 import api
    ext = _register(api, tmp_path)  # noqa: F841 (keeps the library loaded)
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw").replace(*CFG_EDIT))
    assert g.has_feedback()
    raw = np.zeros((96, 128), dtype=np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(128, 96))
    plan = g.plan()
    line = [l for l in plan.splitlines() if "iir_main" in l]
    assert len(line) == 1, plan
    # input, back (the other copy of its own output), output (double buffered)
    assert "[fb:last frame]" in line[0] and "[x2]" in line[0], line[0]
    # the consumer behind it reads this frame's copy, and the chain in front of it is not fused across the module
    assert "b200_pointw (crop+colour)" in plan and plan.count("[fb:last frame]") == 1, plan
    g.close()


def test_feedback_only_module_is_still_traversed(tmp_path):
    """graph-traverse.inc:131-148: a module that is reached through a feedback edge ONLY is run after the rest of the frame."""
    from vkdt_b200 import api
    ext = _register(api, tmp_path)  # noqa: F841
    # iir:02 hangs off colour and feeds nothing but iir:01:back
    cfg = api.DARKROOM_CFG.format(src="i-raw").replace(CFG_EDIT[0],
        "module:iir:01\nmodule:iir:02\nconnect:colour:01:output:iir:01:input\nconnect:colour:01:output:iir:02:input\n"
        "feedback:iir:02:output:iir:02:back\nfeedback:iir:02:output:iir:01:back\nconnect:iir:01:output:filmcurv:01:input\n")
    g = api.Graph(cfg_text=cfg)
    raw = np.zeros((96, 128), dtype=np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(128, 96))
    plan = g.plan()
    lines = [l for l in plan.splitlines() if "iir_main" in l]
    assert len(lines) == 2 and plan.count("[fb:last frame]") == 2, plan
    g.close()


@pytest.mark.gpu
def test_feedback_recurrence_over_frames(gpu, oracle, tmp_path):
    ext = _register(gpu, tmp_path)  # noqa: F841
    w, h = 512, 384
    raw = synth.mosaic(w, h, seed=5)
    WB, CAM = (2.0, 1.0, 1.5), (0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8)
    d = oracle.darkroom_defaults(w, h)
    for k in range(3): d.whitebalance[k] = WB[k]
    for k in range(9): d.cam_to_rec2020[k] = CAM[k]
    col = oracle.darkroom_run(d, raw, stage=5)
    O = oracle

    def rest(img):
        flm, fi = O.new_img(img.shape[0], img.shape[1], 4)
        O.lib().o_filmcurv_main(C.byref(O.img(img)), C.byref(fi), C.byref(d.filmcurv), 1)
        ll, li = O.new_img(img.shape[0], img.shape[1], 4)
        O.lib().o_llap_module(C.byref(O.img(flm)), C.byref(li), C.byref(d.llap), 1)
        want, wi = O.new_img(img.shape[0], img.shape[1], 4)
        O.lib().o_grade_main(C.byref(O.img(ll)), C.byref(wi), C.byref(d.grade), 0)
        return want

    g = gpu.Graph(cfg_text=gpu.DARKROOM_CFG.format(src="i-raw").replace(*CFG_EDIT))
    buf = np.ascontiguousarray(raw)
    g.set_source(buf.ctypes.data, gpu.raw_params(w, h, wb=WB, cam_to_rec2020=CAM))
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    out = np.zeros((oh, ow, 4), dtype=np.float32)
    g.set_sink_buffer(out.ctypes.data, out.nbytes)
    keep = np.float32(0.25)
    back = np.zeros_like(col)
    for frame in range(4):
        g.set_frame(frame)
        g.run(gpu.RUN_RECORD | gpu.RUN_UPLOAD | gpu.RUN_DOWNLOAD | gpu.RUN_WAIT)
        y = col.copy()
        y[..., :3] = ((np.float32(1.0) - keep) * col[..., :3] + keep * back[..., :3]).astype(np.float16).astype(np.float32)
        want = rest(y)
        assert np.array_equal(out[..., :3], want[..., :3]), (frame, float(np.abs(out[..., :3] - want[..., :3]).max()))
        back = y
    # the same frame again (parameters changed, say) reads the same copy: nothing moves
    prev = out.copy()
    g.run(gpu.RUN_RECORD | gpu.RUN_DOWNLOAD | gpu.RUN_WAIT)
    assert np.array_equal(prev, out)
    g.close()
