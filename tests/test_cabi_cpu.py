"""the C-ABI shared library loads without a GPU, exports every symbol include/vkdt_b200.h declares, refuses to
compute without a device (no CPU fallback), and the host-side graph logic (cfg parsing, ROI negotiation, node
creation, fusion, pool layout) behaves like the reference's module layer."""
import os
import re
import numpy as np
import pytest

from vkdt_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "vkdt_b200.h")).read()
    declared = set(re.findall(r"VKB_API[^;(]*?\b(vkb_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(api.DECLARED)
    for s in declared:
        assert hasattr(api.lib, s), s


def test_token_packing_matches_dt_token():
    assert api.token("f16") == 0x363166 and api.token("ui16") == 0x36316975
    assert api.token("demosaic") == int.from_bytes(b"demosaic", "little")
    assert api.token("toolongtoken") == int.from_bytes(b"toolongt", "little")


def test_kernel_registry_covers_the_path():
    k = set(api.kernels())
    for need in [("i-mlv", "unpack"), ("denoise", "noop"), ("hilite", "half"), ("hilite", "reduce"), ("hilite", "assemble"),
                 ("hilite", "doub"), ("demosaic", "gauss"), ("demosaic", "splat"), ("demosaic", "fix"), ("crop", "main"),
                 ("colour", "main"), ("filmcurv", "main"), ("llap", "reduce"), ("llap", "assemble"), ("grade", "main"),
                 ("b200", "pointw"), ("b200", "llapr0"), ("b200", "llapfin"), ("b200", "rawnoop")]:
        assert need in k, need


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.VkbError) as e:
        api.init(0)
    assert e.value.code == -1
    with pytest.raises(api.VkbError):
        api.dispatch("denoise", "noop", [api.image(0, 8, 8, 1, "ui16"), api.image(0, 8, 8, 1, "f16")], b"\0" * 72, b"")
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    raw = np.zeros((64, 64), np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(64, 64))
    with pytest.raises(api.VkbError) as e:
        g.run()
    assert e.value.code == -1


def _plan(src, w, h, extra=(), **kw):
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src=src))
    for l in extra:
        assert g.line(l) == 0, l
    raw = np.zeros((h, w), np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(w, h, **kw))
    return g.plan(), g


@pytest.mark.parametrize("dims", [(6000, 4000), (9504, 6336), (4096, 2160), (512, 384), (402, 410)])
def test_plan_default_darkroom(dims, oracle):
    w, h = dims
    text, g = _plan("i-raw", w, h)
    lines = text.splitlines()
    launches = [l for l in lines if l.startswith("launch")]
    ow, oh = oracle.darkroom_out_size(oracle.darkroom_defaults(w, h))
    assert "sink o-pfm %dx%d" % (ow, oh) in text          # crop's float-math size, same as the oracle's restatement
    assert sum("b200_pointw [b200_pointw (crop+colour+filmcurv)]" in l for l in launches) == 1
    assert sum("b200_llapr0" in l for l in launches) == 1 and sum("llapfin (assemble+colour+grade)" in l for l in launches) == 1
    assert not any("llap_curve" in l or "demosaic_down" in l or "shared_resample" in l for l in launches)
    # the 11 x full-res llap stack is never allocated: no 11-layer image at output resolution
    assert "%dx%dx1x11" % (ow, oh) not in text
    assert text.count("hilite_reduce") == text.count("hilite_assemble")
    # f32 only on the sink edge
    assert sum(":f32@" in l for l in launches) == 1 and launches[-1].rstrip().endswith(tuple("0123456789"))
    pool = int(re.search(r"pool (\d+) bytes", text).group(1))
    # liveness aliasing: far below the sum of all buffers, above the largest live set (mosaic + rgba + f32 out)
    assert pool < 40 * w * h and pool > 16 * ow * oh
    dot = g.dump_nodes()
    for k in ("denoise_noop", "hilite_half", "hilite_doub", "demosaic_down", "demosaic_gauss", "demosaic_splat", "demosaic_fix",
              "shared_resample", "crop_main", "colour_main", "filmcurv_main", "llap_curve", "llap_colour", "grade_main", "o-pfm_main"):
        assert k in dot, k   # the node layer keeps the reference's structure; fusion happens below it


def test_plan_packed_mlv_fuses_unpack_and_noop():
    text, _ = _plan("i-mlv", 4096, 2160, packed_bpp=14)
    assert "b200_rawnoop" in text and "denoise_noop" not in text and "i-mlv_unpack" not in text
    assert "source i-mlv bytes %d packed 14" % (((4096 * 2160 * 14 // 8 + 15) // 16) * 16 + 16) in text


def test_plan_cropped_sensor_keeps_separate_unpack():
    text, _ = _plan("i-mlv", 1024, 768, packed_bpp=14, crop_aabb=(8, 4, 1016, 764))
    assert "i-mlv_unpack" in text and "denoise_noop" in text
    assert "sink o-pfm 1002x754" in text


def test_cfg_params_and_errors():
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    assert g.line("param:filmcurv:01:light:2.5") == 0
    assert g.line("param:filmcurv:01:nosuchparam:1") > 0          # warning, like graph-io.c:52-55
    assert g.line("param:nosuch:01:x:1") > 0
    assert g.line("module:contrast:01") > 0                        # module outside the path: warning, graph still loads
    assert g.line("# comment") == 0
    assert g.line("frames:12") == 0 and g.line("fps:25") == 0
    assert g.line("bogus:1") > 0
    g.line("param:denoise:01:strength:0.4")
    raw = np.zeros((256, 256), np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(256, 256))
    text = g.plan()    # strength > 0 switches the module to the wavelet nodes (denoise/main.c:227-326)
    for k in ("denoise_half", "denoise_downcov", "denoise_down", "denoise_assemble", "denoise_doub"):
        assert k in text, k
    assert text.count("denoise_down [") == 3 and "denoise_noop" not in text


def test_unconnected_graph_is_an_error():
    g = api.Graph(cfg_text="module:i-raw:main\nmodule:display:main\n", sink=None)
    with pytest.raises(api.VkbError):
        g.plan()


def test_plan_rgb_sink_layout():
    """the PFM payload layout: fused producers store r g b themselves, anything else gets one repack launch."""
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    raw = np.zeros((384, 512), np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(512, 384))
    buf = np.zeros(16, np.float32)
    g.set_sink_buffer(buf.ctypes.data, 0)
    text = g.plan()
    fin = [l for l in text.splitlines() if "llapfin" in l][0]
    assert "x4x1:f32@" in fin and "pfmpack" not in text        # memory sinks default to the reference's rgba f32
    g.set_sink_layout(api.SINK_RGB_F32)
    text = g.plan()
    fin = [l for l in text.splitlines() if "llapfin" in l][0]
    assert "x3x1:f32@" in fin and "pfmpack" not in text
    with pytest.raises(api.VkbError):
        g.set_sink_layout(7)


def test_plan_ipfm_source_and_onull_sink(tmp_path):
    """i-pfm uploads f32 and converts once (b200:cvt16); o-null is a sink that keeps the last module's f16 image."""
    from vkdt_b200 import synth
    fn = str(tmp_path / "x.pfm")
    synth.write_pfm(fn, np.random.default_rng(0).random((40, 64, 3), dtype=np.float32))
    cfg = ("module:i-pfm:main\nmodule:filmcurv:01\nmodule:o-null:main\n"
           "connect:i-pfm:main:output:filmcurv:01:input\nconnect:filmcurv:01:output:o-null:main:input\n"
           "param:i-pfm:main:filename:%s\n" % fn)
    g = api.Graph(cfg_text=cfg, sink=None)
    text = g.plan()
    assert "b200_cvt16" in text and "filmcurv_main" in text and "sink o-null 64x40" in text
    assert "64x40x4x1:f32@" in text and "pfmpack" not in text
    mono = str(tmp_path / "y.pfm")
    synth.write_pfm(mono, np.zeros((8, 12), np.float32))
    g = api.Graph(cfg_text=cfg.replace(fn, mono), sink=None)
    # single channel ("Pf") files upload one float per pixel (the chain's kernels then refuse the 1-channel image at launch)
    assert "12x8x1x1:f32@" in g.plan()


@pytest.mark.parametrize("crop,rot", [((0.25, 0.75, 0.25, 0.75), 0.0), ((0.1, 0.9, 0.2, 0.7), 0.0), ((0.0, 1.0, 0.0, 1.0), 90.0),
                                      ((0.05, 0.95, 0.05, 0.95), 7.5), ((1.0, 3.0, 3.0, 7.0), 1337.0)])
def test_crop_roi_follows_the_reference_arithmetic(oracle, crop, rot):
    """the sink size falls out of crop's float math (crop/main.c:228-275): same numbers as the oracle's restatement for
    explicit crop windows, a quarter turn, a free rotation and the magic defaults; paramsub edits one element."""
    w, h = 1200, 802
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    assert g.line("param:crop:01:crop:%g:%g:%g:%g" % crop) == 0
    assert g.line("param:crop:01:rotate:%g" % rot) == 0
    raw = np.zeros((h, w), np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(w, h))
    g.set_sink_buffer(np.zeros(4, np.float32).ctypes.data, 0)
    text = g.plan()
    d = oracle.darkroom_defaults(w, h)
    for k in range(4):
        d.crop.crop[k] = crop[k]
    d.crop.rotate = rot
    ow, oh = oracle.darkroom_out_size(d)
    assert "sink o-pfm %dx%d" % (ow, oh) in text, (ow, oh, [l for l in text.splitlines() if l.startswith("sink")])
    if rot == 0.0:
        assert g.line("paramsub:crop:01:crop:1:2:%g" % 0.5) == 0    # module:inst:param:beg:end:values: element 1 (right edge) only
        d.crop.crop[1] = 0.5
        ow2, oh2 = oracle.darkroom_out_size(d)
        assert "sink o-pfm %dx%d" % (ow2, oh2) in g.plan() and oh2 == oh and ow2 != ow
        assert g.line("param:crop:01:crop:0.5:0.5:0.2:0.7") == 0      # a window without area must not plan
        with pytest.raises(api.VkbError):
            g.plan()


def test_cyclic_connections_are_refused_like_the_reference():
    """a connection that would close a loop is refused with code 12 before anything is touched (connector.c:21-27, cycles.h:47-66):
    a warning for the config reader, and the graph stays what it was."""
    cfg = ("module:i-raw:main\nmodule:crop:01\nmodule:colour:01\nmodule:filmcurv:01\nmodule:display:main\n"
           "connect:i-raw:main:output:crop:01:input\nconnect:crop:01:output:colour:01:input\nconnect:colour:01:output:filmcurv:01:input\n"
           "connect:filmcurv:01:output:display:main:input\n")
    g = api.Graph(cfg_text=cfg)
    raw = np.zeros((64, 96), np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(96, 64))
    before = g.plan()
    assert g.line("connect:filmcurv:01:output:colour:01:input") == 12      # colour <- filmcurv <- colour
    assert g.line("connect:colour:01:output:colour:01:input") == 12        # onto itself
    assert g.line("connect:filmcurv:01:output:crop:01:input") == 12        # two modules upstream
    assert g.plan() == before


def test_plan_from_mlv_files(tmp_path):
    """the container reader on well-formed clips: uncompressed -> packed upload + device unpack, lossless -> u16 upload."""
    from vkdt_b200 import synth
    yy, xx = np.mgrid[0:66, 0:128]
    pix = (2048 + 31 * xx + 17 * yy).astype(np.uint16)
    for lossless, want in ((False, "packed 14"), (True, "packed 0")):
        fn = str(tmp_path / ("l.mlv" if lossless else "u.mlv"))
        synth.write_mlv(fn, [pix, pix, pix], bpp=14, lossless=lossless, camera_name="Canon EOS 5D Mark III")
        g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-mlv"))
        assert g.line("param:i-mlv:main:filename:%s" % fn) == 0
        text = g.plan()
        assert want in text and "sink o-pfm 128x66" in text, text[-300:]   # no micro-crop below 400 px (crop/main.c:194-224)


def test_cli_dump_nodes_without_a_gpu(tmp_path):
    """`vkdt-b200-cli -g x.cfg --dump-nodes` (graph-print.h:76) only needs the host half: cfg file next to a dng, searchpath
    resolution of the relative filename, node layer as graphviz."""
    import os
    import subprocess
    from vkdt_b200 import synth
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "vkdt_b200", "vkdt-b200-cli")
    if not os.path.exists(cli):
        pytest.skip("cli not built")
    synth.write_dng(str(tmp_path / "img.dng"), np.zeros((402, 600), np.uint16))
    cfg = tmp_path / "img.dng.cfg"
    cfg.write_text(api.DARKROOM_CFG.format(src="i-raw") + "param:i-raw:main:filename:img.dng\n")
    r = subprocess.run([cli, "-g", str(cfg), "--dump-nodes"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    # the reference's default export: o-jpg behind a colenc (cli/main.c:58-59, graph-export.c:66-86)
    assert r.stdout.lstrip().startswith("digraph") and "hilite_reduce" in r.stdout and "llap_curve" in r.stdout and "o-jpg_main" in r.stdout and "colenc_main" in r.stdout
    r = subprocess.run([cli, "-g", str(cfg), "--dump-nodes", "--format", "o-pfm", "--colour-prim", "bt2020", "--colour-trc", "linear"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "o-pfm_main" in r.stdout and "colenc" not in r.stdout
    # --width / --height (cli/main.c:68-71): a resize module in front of the sink; beyond a factor of three behind a separable blur
    r = subprocess.run([cli, "-g", str(cfg), "--dump-nodes", "--width", "250"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "resize_main" in r.stdout and "shared_blurh" not in r.stdout and "colenc_main" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([cli, "-g", str(cfg), "--dump-nodes", "--width", "100", "--height", "100"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "resize_main" in r.stdout and "shared_blurh" in r.stdout and "shared_blurv" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([cli, "-g", str(tmp_path / "missing.cfg"), "--dump-nodes"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    r = subprocess.run([cli, "--bogus"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "usage" in r.stderr


def test_jpeg_writer_round_trip(tmp_path):
    """the baseline jpeg writer behind o-jpg (host only): a decoder reads it back close to what went in, at about the size
    and fidelity of libjpeg at the same quality."""
    import ctypes as C
    import io
    import numpy as np
    from PIL import Image
    from vkdt_b200 import api
    w, h = 333, 201                                   # not a multiple of 8: edge blocks
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, 4), np.uint8)
    img[..., 0] = xx * 255 // w; img[..., 1] = yy * 255 // h
    img[..., 2] = ((np.sin(xx / 7.0) * np.cos(yy / 5.0) * 0.5 + 0.5) * 255).astype(np.uint8); img[..., 3] = 255
    img[50:90, 100:180, :3] = [250, 10, 30]
    fn = str(tmp_path / "t.jpg")
    for q in (95, 60):
        api.check(api.lib.vkb_jpeg_write(fn.encode(), img.ctypes.data_as(C.c_void_p), w, h, float(q)))
        dec = np.asarray(Image.open(fn).convert("RGB")).astype(float)
        ref = io.BytesIO()
        Image.fromarray(img[..., :3]).save(ref, "JPEG", quality=q, subsampling=0)
        pil = np.asarray(Image.open(io.BytesIO(ref.getvalue())).convert("RGB")).astype(float)
        ps = lambda a: 10 * np.log10(255.0 ** 2 / ((a - img[..., :3]) ** 2).mean())
        assert dec.shape == (h, w, 3) and ps(dec) > ps(pil) - 0.5 and os.path.getsize(fn) < 1.1 * len(ref.getvalue())


def test_export_inserts_colenc_like_the_reference():
    """graph-export.c:66-86: an 8 bit sink (o-jpg) or a colour space other than linear bt2020 puts colenc in front of the sink;
    the launch plan fuses it behind the pointwise chain when there is no llap, else runs it as the last launch."""
    import numpy as np
    from vkdt_b200 import api
    raw = np.zeros((384, 512), dtype=np.uint16)
    for sink, prim, trc, want in (("o-jpg", 1, 1, True), ("o-pfm", 2, 0, False), ("o-pfm", 4, 2, True)):
        g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"), sink=sink, prim=prim, trc=trc)
        g.set_source(raw.ctypes.data, api.raw_params(512, 384))
        plan = g.plan()
        assert ("colenc" in plan) == want, (sink, prim, trc, plan)
        if sink == "o-jpg":
            assert "ui8" in plan
        g.close()
    cfg = api.DARKROOM_CFG.format(src="i-raw").replace("connect:filmcurv:01:output:llap:01:input\n", "").replace(
        "connect:llap:01:output:grade:01:input\n", "connect:filmcurv:01:output:grade:01:input\n")
    g = api.Graph(cfg_text=cfg, sink="o-jpg", prim=1, trc=1)
    g.set_source(raw.ctypes.data, api.raw_params(512, 384))
    assert "crop+colour+filmcurv+grade+colenc" in g.plan()
    g.close()
