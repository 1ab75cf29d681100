"""export side: colenc (colenc/main.comp) and the 8 bit sinks.  vkdt-cli's default export is o-jpg in
sRGB primaries with the rec709 curve (cli/main.c:58-59), i.e. every frame ends in colenc writing an rgba:ui8 image
(graph-export.c:66-86).  kernel against the oracle for every primaries / curve pair, the whole graph into 8 bit memory sinks
(rgba, the reference's mapped image, and packed rgb) with exact integer agreement, and the cli writing a jpg."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from helpers import to_dev_f16, dev_f16, dev_f32, to_host, ibits, f16_ulp_diff, parity_gate
from vkdt_b200 import synth

pytestmark = pytest.mark.gpu
WB = (2.0, 1.0, 1.5)
CAM = (0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8)


def _rgb(w, h, seed=3):
    rng = np.random.default_rng(seed)
    a = np.zeros((h, w, 4), dtype=np.float32)
    a[..., :3] = rng.random((h, w, 3), dtype=np.float32) ** 2.2 * 1.2      # some values above 1
    a[::7, ::5, :3] = 0.0
    a[3::11, 2::13, 0] = -0.01                                              # a few negative (out of gamut) values
    a[..., 3] = 1.0
    return a.astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("prim", [1, 2, 3, 4, 5, 6, 7, 10])
@pytest.mark.parametrize("trc", [0, 1, 2, 3, 4, 5, 6])
def test_colenc_kernel(gpu, oracle, prim, trc):
    import torch
    O = oracle
    w, h = 160, 96
    a = _rgb(w, h)
    if trc in (3, 4, 6):
        a = np.maximum(a, 0.0)                    # pow of a negative base: undefined in the shader, NaN either way
    O.lib().o_colenc_main.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    want, wi = O.new_img(h, w, 4)
    O.lib().o_colenc_main(C.byref(O.img(a)), C.byref(wi), prim, trc, 1)
    d_out = dev_f16(h, w, 4)
    gpu.dispatch("colenc", "main", [gpu.image(to_dev_f16(a), w, h, 4, "f16"), gpu.image(d_out, w, h, 4, "f16")], b"", ibits(prim, trc))
    got = to_host(d_out)
    d = f16_ulp_diff(got[..., :3], want[..., :3])
    assert d.max() == 0, "prim %d trc %d: max ulp %d, %d values differ" % (prim, trc, int(d.max()), int((d != 0).sum()))
    # 8 bit store
    want8, wi8 = O.new_img(h, w, 4)
    O.lib().o_colenc_main(C.byref(O.img(a)), C.byref(wi8), prim, trc, 2)
    d8 = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    gpu.dispatch("colenc", "main", [gpu.image(to_dev_f16(a), w, h, 4, "f16"), gpu.image(d8, w, h, 4, "ui8")], b"", ibits(prim, trc))
    assert np.array_equal(d8.cpu().numpy().astype(np.float32), want8)


def _graph8(gpu, raw, layout, strength=0.0, prim=1, trc=1):
    h, w = raw.shape
    g = gpu.Graph(cfg_text=gpu.DARKROOM_CFG.format(src="i-raw"), sink="o-jpg", prim=prim, trc=trc)
    if strength > 0:
        assert g.line("param:denoise:01:strength:%g" % strength) == 0
    buf = np.ascontiguousarray(raw)
    g.set_source(buf.ctypes.data, gpu.raw_params(w, h, wb=WB, cam_to_rec2020=CAM, noise_a=100.0, noise_b=2.0))
    g.set_sink_layout(layout)
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    out = np.zeros((oh, ow, 4 if layout == gpu.SINK_RGBA_UI8 else 3), dtype=np.uint8)
    g.set_sink_buffer(out.ctypes.data, out.nbytes)
    g.run(gpu.RUN_RECORD | gpu.RUN_UPLOAD | gpu.RUN_DOWNLOAD | gpu.RUN_WAIT)
    plan = g.perf() or g.plan()
    g.close()
    return out, plan


def _oracle8(oracle, raw, strength=0.0, prim=1, trc=1):
    h, w = raw.shape
    d = oracle.darkroom_defaults(w, h)
    for k in range(3): d.whitebalance[k] = WB[k]
    for k in range(9): d.cam_to_rec2020[k] = CAM[k]
    d.noise_a, d.noise_b = 100.0, 2.0
    d.denoise.strength = strength
    d.enable_colenc, d.colenc_prim, d.colenc_trc, d.sink_unorm8 = 1, prim, trc, 1
    return oracle.darkroom_run(d, raw)


@pytest.mark.parametrize("strength", [0.0, 0.4])
def test_default_export_into_8bit_sinks(gpu, oracle, strength):
    """what `vkdt-cli -g x.cfg` develops (o-jpg, sRGB primaries, rec709 curve): the 8 bit image the encoder is handed."""
    w, h = 644, 486
    raw = synth.mosaic(w, h, seed=91)
    want = _oracle8(oracle, raw, strength)
    rgba, _ = _graph8(gpu, raw, gpu.SINK_RGBA_UI8, strength)
    assert rgba.shape == want.shape
    assert np.array_equal(rgba.astype(np.float32), want), int((rgba.astype(np.float32) != want).sum())
    rgb, _ = _graph8(gpu, raw, gpu.SINK_RGB_UI8, strength)
    assert np.array_equal(rgb, rgba[..., :3])


def test_f32_export_in_another_colour_space(gpu, oracle):
    """o-pfm with --colour-prim P3 --colour-trc sRGB: colenc in front of the f32 sink."""
    w, h = 512, 384
    raw = synth.mosaic(w, h, seed=17)
    d = oracle.darkroom_defaults(w, h)
    for k in range(3): d.whitebalance[k] = WB[k]
    for k in range(9): d.cam_to_rec2020[k] = CAM[k]
    d.enable_colenc, d.colenc_prim, d.colenc_trc = 1, 4, 2
    want = oracle.darkroom_run(d, raw)
    g = gpu.Graph(cfg_text=gpu.DARKROOM_CFG.format(src="i-raw"), sink="o-pfm", prim=4, trc=2)
    buf = np.ascontiguousarray(raw)
    g.set_source(buf.ctypes.data, gpu.raw_params(w, h, wb=WB, cam_to_rec2020=CAM))
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    out = np.zeros((oh, ow, 4), dtype=np.float32)
    g.set_sink_buffer(out.ctypes.data, out.nbytes)
    g.run(gpu.RUN_RECORD | gpu.RUN_UPLOAD | gpu.RUN_DOWNLOAD | gpu.RUN_WAIT)
    parity_gate(out[..., :3], want[..., :3], "o-pfm in P3 / sRGB curve")


def test_cli_writes_a_jpg(gpu, oracle, tmp_path):
    """the reference's default invocation: no --format, no colour flags -> <filename>.jpg in sRGB / rec709 curve."""
    from PIL import Image
    w, h = 644, 486
    raw = synth.mosaic(w, h, seed=91)
    dng = tmp_path / "frame.dng"
    synth.write_dng(str(dng), raw, black=2048, white=15000)
    cfg = tmp_path / "frame.dng.cfg"
    cfg.write_text(gpu.DARKROOM_CFG.format(src="i-raw") + "param:i-raw:main:filename:%s\n" % dng)
    cli = os.path.join(os.path.dirname(gpu.LIB_PATH), "vkdt-b200-cli")
    out = str(tmp_path / "out")
    r = subprocess.run([cli, "-g", str(cfg), "--filename", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    img = np.asarray(Image.open(out + ".jpg").convert("RGB")).astype(np.float64)
    assert img.shape[0] > 400 and img.shape[1] > 600 and img.std() > 10.0
