"""regenerates the committed golden vectors.  run in the build container (needs /root/reference for the MLV part):

    python tests/golden/make_golden.py

 - mlv_unpack_{10,12,14}.npz: seeded pixels packed into an MLV file and decoded by the REFERENCE's own
   video_mlv.c (oracle/_ref/libmlvref.so).  pins the oracle's o_mlv_unpack and the CUDA unpack kernel bit exactly.
 - host_ref.npz: crop / colour parameter blocks from the REFERENCE's own crop/main.c and colour/main.c (oracle/_ref/libhostref.so).
 - host_nodes.json.gz: node graphs of denoise / hilite / demosaic / llap / filmcurv from the REFERENCE's own create_nodes (same library).
 - host_graph.json.gz: the whole module pass (config reader, roi negotiation, nodes, committed parameters) of the REFERENCE's own
   graph code over its own default darkroom config (same library).
 - host_cfg.json.gz: return codes and effects of config lines from the REFERENCE's own graph-io.c (same library).
 - shader_ref.npz / shader_ref.json: outputs (images for the filtering kernels, sha256 digests for the bit exact ones) of the REFERENCE's own compute shaders compiled as C++ (oracle/glsl -> oracle/_ref/libshaderref.so).
 - pipeline_ref.npz: the default darkroom graph end to end from the REFERENCE's own graph code + own shaders on the CPU.
 - darkroom_*.npz: outputs of the CPU oracle for the default darkroom graph on a small synthetic frame.  these pin
   the oracle's whole graph against accidental edits (the reference-made vectors for the float path are shader_ref.* and pipeline_ref.npz).
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from vkdt_b200 import synth          # noqa: E402
from oracle import oracle_py as O    # noqa: E402


def mlv_goldens():
    O.build()
    for bpp in (10, 12, 14):
        rng = np.random.default_rng(0xC0FFEE + bpp)
        w, h = 136, 50
        pix = rng.integers(0, 1 << bpp, (h, w), dtype=np.uint16)
        pix[0, :4] = [0, (1 << bpp) - 1, 1, (1 << bpp) - 2]
        with tempfile.TemporaryDirectory() as td:
            fn = os.path.join(td, "g.mlv")
            synth.write_mlv(fn, [pix], bpp=bpp, black=64, white=(1 << bpp) - 1)
            out, info = O.ref_mlv_decode(fn, 0)
            payload = open(fn, "rb").read()
        assert info["bpp"] == bpp and (out == pix).all(), "reference decoder disagrees with the packer"
        words = synth.pack_bits(pix, bpp)
        np.savez_compressed(os.path.join(HERE, "mlv_unpack_%d.npz" % bpp), words=words, expected=out, width=w, height=h, bpp=bpp)
        print("mlv golden", bpp, out.shape, len(payload))


def lj92_goldens():
    """lossless jpeg: (1) a stream written by the REFERENCE's own encoder (liblj92 lj92_encode) and decoded by its decoder,
    (2) streams of our test encoder (all predictors, 1 and 2 components) decoded by the reference's decoder, (3) a lossless
    MLV clip read through the reference's mlv_get_frame.  pins vkb_lj92_decode / the i-mlv lossless path bit exactly."""
    import ctypes as C
    R = O.ref_lib()
    assert R is not None, "oracle/_ref/libmlvref.so missing: run `make -C oracle ref` where /root/reference exists"
    R.lj92_open.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 4
    R.lj92_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    R.lj92_close.argtypes = [C.c_void_p]
    R.lj92_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]

    def ref_decode(stream):
        buf = (C.c_uint8 * len(stream)).from_buffer_copy(stream)
        hdl = C.c_void_p()
        w, h, b, c = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        assert R.lj92_open(C.byref(hdl), buf, len(stream), C.byref(w), C.byref(h), C.byref(b), C.byref(c)) == 0
        out = np.zeros((h.value, w.value * c.value), dtype=np.uint16)
        assert R.lj92_decode(hdl, out.ctypes.data_as(C.c_void_p), out.size, 0, None, 0) == 0
        R.lj92_close(hdl)
        return out, b.value, c.value

    rng = np.random.default_rng(0x1A92)
    streams, expected, notes = [], [], []
    # (1) reference encoder
    w, h, bits = 96, 40, 14
    img = (synth.mosaic(w, h, seed=5) & 0x3fff).astype(np.uint16)
    enc, n = C.c_void_p(), C.c_int()
    assert R.lj92_encode(img.ctypes.data_as(C.c_void_p), w, h, bits, w, 0, None, 0, C.byref(enc), C.byref(n)) == 0
    s = C.string_at(enc, n.value)
    out, b, c = ref_decode(s)
    assert (out == img).all() and b == bits
    streams.append(np.frombuffer(s, np.uint8)); expected.append(out); notes.append("reference encoder, %d bit" % bits)
    # (2) our test encoder, decoded by the reference.  multi-component streams only with predictor 1 (what Canon's encoder
    # writes): for predictors 2..7 the vendored liblj92's row loops assume one component and deviate from T.81
    for bits, comps, pred, lengths in [(14, 1, 1, None), (14, 2, 1, None), (12, 1, 4, None), (14, 2, 1, [2, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12]),
                                       (14, 1, 6, [2, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12]),
                                       (14, 1, 7, None), (10, 1, 5, None), (14, 1, 2, None), (14, 1, 3, None)]:
        img = rng.integers(0, 1 << bits, (20, 36)).astype(np.uint16)
        img[2:9, 4:30] = (np.arange(26)[None, :] * 5 + 300) & ((1 << bits) - 1)
        s = synth.lj92_encode(img, bits, comps, pred, lengths)
        out, b, c = ref_decode(s)
        assert (out == img).all() and b == bits and c == comps, (bits, comps, pred)
        streams.append(np.frombuffer(s, np.uint8)); expected.append(out); notes.append("test encoder bits %d comps %d predictor %d" % (bits, comps, pred))
    # (3) lossless clip through the reference's mlv reader
    yy, xx = np.mgrid[0:34, 0:64]
    pix = (2048 + 37 * xx + 11 * yy + rng.integers(0, 16, (34, 64))).astype(np.uint16)   # smooth: compresses below the packed size
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "l.mlv")
        synth.write_mlv(fn, [pix, pix[::-1].copy()], bpp=14, lossless=True)
        f0, info = O.ref_mlv_decode(fn, 0)
        f1, _ = O.ref_mlv_decode(fn, 1)
        clip = np.frombuffer(open(fn, "rb").read(), np.uint8)
    assert (f0 == pix).all() and (f1 == pix[::-1]).all()
    np.savez_compressed(os.path.join(HERE, "lj92.npz"), clip=clip, clip_frames=np.stack([f0, f1]), notes=np.array(notes),
                        **{"stream_%d" % i: s for i, s in enumerate(streams)}, **{"expected_%d" % i: e for i, e in enumerate(expected)})
    print("lj92 goldens:", len(streams), "streams + lossless clip", clip.size, "bytes")


def host_goldens():
    """crop and colour parameter blocks from the REFERENCE's own crop/main.c and colour/main.c (host side, compiled in place:
    oracle/_ref/libhostref.so).  pins o_crop_roi_out / o_crop_commit / o_colour_commit and the product's module callbacks."""
    assert O.ref_host_lib() is not None, "oracle/_ref/libhostref.so missing: run `make -C oracle ref` where /root/reference exists"
    rng = np.random.default_rng(0xC0107)
    crop_in, crop_out = [], []
    for t in range(300):
        w, h = int(rng.integers(20, 9000)), int(rng.integers(20, 9000))
        ori = int(rng.choice([0, 1, 3, 6, 8]))
        if t % 3 == 0:
            crop = (1.0, 3.0, 3.0, 7.0)
        else:
            a, b = sorted(rng.random(2)); c, d = sorted(rng.random(2))
            crop = (float(a), float(b) + 0.01, float(c), float(d) + 0.01)
        rot = float(rng.choice([1337.0, 0.0, 90.0, 180.0, 270.0, float(rng.uniform(-30, 30))]))
        persp = np.array([0.25, 0.25, 0.75, 0.25, 0.75, 0.75, 0.25, 0.75]) + (0 if t % 4 == 0 else rng.uniform(-0.05, 0.05, 8))
        inp = np.array([ori, w, h, *persp, *crop, rot], dtype=np.float64)
        ow, oh, f = O.ref_crop(ori, w, h, [float(np.float32(x)) for x in persp], crop, rot)
        o2 = O.crop_oracle(ori, w, h, [float(np.float32(x)) for x in persp], crop, rot)
        assert (ow, oh) == o2[:2] and np.array_equal(f.view(np.uint32), o2[2].view(np.uint32)), "oracle crop differs from the reference"
        crop_in.append(inp); crop_out.append(np.concatenate([[ow, oh], f.astype(np.float64)]))
    col_par, col_img, col_out = [], [], []
    for t in range(120):
        d = O.darkroom_defaults(64, 64)
        p = d.colour
        p.exposure = float(rng.uniform(-2, 2)); p.sat = float(rng.choice([1.0, rng.uniform(0, 2)]))
        p.matrix = int(rng.choice([0, 1, 2, 3, 4, 5])); p.gamut = int(rng.integers(0, 3)); p.clip = int(rng.integers(0, 2)); p.clipmax = float(rng.uniform(0.5, 2))
        p.temp = float(rng.choice([6504.0, 0.0, rng.uniform(2000, 10000)])); p.picked = int(rng.integers(0, 2))
        for k in range(9): p.mat[k] = float(rng.uniform(-1, 1.5))
        for k in range(4): p.white[k] = float(rng.choice([0.0, rng.uniform(0.2, 1.0)]))
        p.mode = int(rng.integers(0, 2)); p.cnt = int(rng.integers(0, 25))
        for k in range(144): p.rbmap[k] = float(rng.uniform(0.05, 0.9))
        wb = [float(np.float32(rng.uniform(0.5, 3))), 1.0, float(np.float32(rng.uniform(0.5, 3))), 1.0]
        cam = [float(np.float32(x)) for x in (np.eye(3) + rng.uniform(-0.3, 0.3, (3, 3))).ravel()]
        prim, trc = (0, 0) if t < 80 else (int(rng.choice([1, 2, 3, 4, 5, 6, 7])), int(rng.choice([0, 1, 2, 3, 4, 5, 6])))
        raw = bytes(p)
        f, wbo = O.ref_colour_commit(raw, wb, cam, prim, trc)
        f2, wbo2 = O.colour_commit_oracle(raw, wb, cam, prim, trc)
        assert np.array_equal(np.nan_to_num(f, nan=-7.0), np.nan_to_num(f2, nan=-7.0)) and np.array_equal(wbo, wbo2), "oracle colour commit differs from the reference"
        col_par.append(np.frombuffer(raw, np.uint8)); col_img.append(np.array(wb + cam + [prim, trc], np.float64)); col_out.append(np.concatenate([f, wbo]))
    np.savez_compressed(os.path.join(HERE, "host_ref.npz"), crop_in=np.array(crop_in), crop_out=np.array(crop_out),
                        colour_params=np.array(col_par), colour_img=np.array(col_img), colour_out=np.array(col_out, dtype=np.float32))
    print("host goldens:", len(crop_in), "crop cases,", len(col_par), "colour cases")


NODE_CASES = [  # (cfg lines, width, height, raw parameters)
    ([], 512, 384, {}),
    (["param:denoise:01:strength:0.4"], 512, 384, dict(wb=(2.0, 1.0, 1.5), noise_a=100.0, noise_b=2.0)),
    (["param:denoise:01:strength:0.4"], 516, 390, dict(filters=9, wb=(2.0, 1.0, 1.5), noise_a=100.0, noise_b=2.0)),
    ([], 516, 390, dict(filters=9)),
    (["param:demosaic:01:method:1"], 512, 384, {}),
    (["param:demosaic:01:method:2"], 512, 384, {}),
    (["param:demosaic:01:method:2"], 516, 390, dict(filters=9)),
    (["param:denoise:01:strength:0.4"], 1030, 778, dict(crop_aabb=(8, 4, 1000, 700), black=1024.0, white=16000.0)),
    ([], 1030, 778, dict(crop_aabb=(8, 4, 1000, 700), black=1024.0, white=16000.0)),
    ([], 6000, 4000, {}),
    (["param:denoise:01:strength:0.4", "param:demosaic:01:method:1"], 9504, 6336, dict(wb=(2.0, 1.0, 1.5))),
    (["param:denoise:01:strength:0.4"], 6240, 4152, dict(filters=9, wb=(2.0, 1.0, 1.5))),
    ([], 4096, 2160, dict(wb=(2.0, 1.0, 1.5))),
    (["param:denoise:01:strength:0.4"], 16384, 12288, {}),
    ([], 38, 26, {}),
    (["param:denoise:01:strength:0.2"], 70, 50, {}),
]
NODE_MODULES = ("denoise", "hilite", "demosaic", "llap", "filmcurv")


def reference_nodes(lines, w, h, kw):
    """{module: text} from the REFERENCE's own create_nodes (oracle/_ref/libhostref.so) for the modules of the default darkroom
    graph that build nodes, fed with what the product's vkb_graph_describe says enters each module."""
    from vkdt_b200 import api
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    for ln in lines:
        assert g.line(ln) == 0, ln
    raw = np.zeros((h, w), np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(w, h, **kw))
    blocks = O.parse_described_modules(g.describe())
    names, img, out = list(blocks), None, {}
    for i, m in enumerate(names):
        if m not in NODE_MODULES:
            img = None
            continue
        out[m], img = O.ref_nodes(m, blocks[names[i - 1]], blocks[m], img_line=img)
    return out


def node_goldens():
    """the node graphs of denoise / hilite / demosaic / llap / filmcurv as the REFERENCE's own <module>/main.c builds them (roi
    callbacks + create_nodes compiled in place, oracle/ref_nodes_shim.c): node and kernel names, dispatch sizes, push constants,
    connector channels / formats / sizes and the wiring.  pins the product's module callbacks (tests/test_host_ref_cpu.py)."""
    import json
    assert O.ref_host_lib() is not None, "oracle/_ref/libhostref.so missing: run `make -C oracle ref` where /root/reference exists"
    cases = [dict(lines=ln, w=w, h=h, raw=kw, modules=reference_nodes(ln, w, h, kw)) for ln, w, h, kw in NODE_CASES]
    import gzip
    with gzip.GzipFile(os.path.join(HERE, "host_nodes.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(cases, indent=0).encode())
    print("node goldens:", len(cases), "graphs,", sum(len(c["modules"]) for c in cases), "module node lists")


GRAPH_CASES = NODE_CASES + [  # parameter lines travel through the reference's own config reader (graph-io.c) here
    (["param:crop:01:rotate:7.5", "param:crop:01:crop:0.1:0.9:0.2:0.8"], 600, 400, {}),
    (["param:crop:01:rotate:90"], 600, 400, dict(wb=(2.0, 1.0, 1.5))),
    (["param:crop:01:perspect:0.2:0.25:0.8:0.2:0.75:0.8:0.25:0.75", "param:denoise:01:strength:0.3"], 804, 602, {}),
    (["param:colour:01:exposure:0.75", "param:colour:01:sat:1.25", "param:colour:01:matrix:1", "param:colour:01:temp:4500"], 512, 384, dict(wb=(2.2, 1.0, 1.4))),
    (["param:colour:01:matrix:2", "param:colour:01:mat:1.1:-0.05:-0.05:0.02:0.9:0.08:0.0:-0.1:1.1", "param:colour:01:clip:1"], 512, 384, dict(wb=(1.9, 1.0, 1.7))),
    (["param:filmcurv:01:light:2.5", "param:filmcurv:01:contrast:1.4", "param:filmcurv:01:colour:1", "param:llap:01:clarity:0.5", "param:llap:01:sigma:0.2",
      "param:grade:01:gain:1.1:1.0:0.9:1.0", "param:hilite:01:white:0.9", "param:denoise:01:luma:0.3", "param:demosaic:01:colour:1"], 516, 390, dict(filters=9)),
    # "#export:max:<w>:<h>": vkdt-cli --width / --height, i.e. dt_graph_replace_display with a resize module (blur + slice above a
    # factor of three, catmull-rom below, bypass when nothing shrinks)
    (["#export:max:200:200"], 640, 480, {}),
    (["#export:max:400:0", "param:denoise:01:strength:0.3"], 640, 480, dict(wb=(2.0, 1.0, 1.5))),
    (["#export:max:0:500"], 1002, 666, dict(filters=9)),
    (["#export:max:1000:1000"], 322, 246, {}),
    # "#export:colour:<prim>:<trc>" / "#export:sink:<module>": --colour-prim / --colour-trc / --format, i.e. the colenc module that
    # dt_graph_replace_display puts in front of an 8 bit sink or a colour space other than linear bt2020 (graph-export.c:66-86)
    (["#export:colour:1:1", "#export:sink:o-jpg"], 640, 480, {}),                                  # what `vkdt-cli -g x.cfg` does
    (["#export:colour:4:2"], 640, 480, dict(wb=(2.0, 1.0, 1.5))),                                  # f32 pfm in display p3 with the srgb curve
    (["#export:colour:1:1", "#export:sink:o-jpg", "#export:max:300:0", "param:denoise:01:strength:0.3"], 804, 602, {}),
    (["#export:colour:2:0", "#export:sink:o-jpg"], 322, 246, {}),                                  # 8 bit linear bt2020: colenc for the format alone
    # feedback edges through the module pass (second traversal round, frames = 2)
    (["feedback:grade:01:output:colour:01:spectra"], 640, 480, {}),
    (["feedback:llap:01:output:colour:01:spectra", "param:denoise:01:strength:0.3"], 1002, 668, {}),
]


MLV_CLIPS = {  # name: (width, height, bits, black, white, camera name in IDNT or None); cameras outside dcraw's adobe_coeff table
    "plain14": (512, 384, 14, 2048, 15000, None),
    "named12": (328, 246, 12, 512, 3900, "Synth Cam One"),
}
MLV_DIR = "/tmp/vkdt_b200_golden"   # the file name is part of i-mlv's parameter block: the tests write the clips to the same place


def write_golden_clip(name):
    w, h, bpp, black, white, camera = MLV_CLIPS[name]
    os.makedirs(MLV_DIR, exist_ok=True)
    fn = os.path.join(MLV_DIR, name + ".mlv")
    pix = (synth.mosaic(w, h, seed=5) >> (14 - bpp)).astype(np.uint16)
    synth.write_mlv(fn, [pix, pix[::-1].copy()], bpp=bpp, black=black, white=white, camera_name=camera)
    return fn


PFM_CFG = ("module:i-pfm:main\nmodule:colour:01\nmodule:filmcurv:01\nmodule:display:main\nconnect:i-pfm:main:output:colour:01:input\n"
           "connect:colour:01:output:filmcurv:01:input\nconnect:filmcurv:01:output:display:main:input\nparam:i-pfm:main:filename:%s\n")


def write_golden_pfm():
    """an intermediate image as another vkdt would hand it over with o-pfm, and a cfg that develops it from colour on"""
    os.makedirs(MLV_DIR, exist_ok=True)
    fn = os.path.join(MLV_DIR, "mid.pfm")
    synth.write_pfm(fn, np.random.default_rng(0).random((40, 64, 3), dtype=np.float32))
    cfg = os.path.join(MLV_DIR, "pfm.cfg")
    with open(cfg, "w") as f:
        f.write(PFM_CFG % fn)
    return fn, cfg


LUT_LINES = ["module:i-lut:abney", "module:i-lut:spectra", "param:i-lut:abney:filename:%(dir)s/abney.lut", "param:i-lut:spectra:filename:%(dir)s/spectra.lut",
             "connect:i-lut:abney:output:colour:01:abney", "connect:i-lut:spectra:output:colour:01:spectra", "param:colour:01:gamut:2", "param:colour:01:sat:1.2"]


CLUT_LINES = ["module:i-lut:clut", "param:i-lut:clut:filename:%(dir)s/clut.lut", "connect:i-lut:clut:output:colour:01:clut", "param:colour:01:matrix:4", "param:colour:01:temp:5200"]


def write_golden_luts(clut=False):
    """abney (rg f16) and spectra (rgba f32) tables of the shapes the reference's tools write (core/lut.h), synthetic content"""
    import struct
    os.makedirs(MLV_DIR, exist_ok=True)
    rng = np.random.default_rng(4)
    for name, a in (("abney", rng.uniform(0.25, 0.4, (40, 64, 2)).astype(np.float16)), ("spectra", rng.uniform(0.1, 1.0, (48, 48, 4)).astype(np.float32))):
        with open(os.path.join(MLV_DIR, name + ".lut"), "wb") as f:
            f.write(struct.pack("<IHBBII", 1234, 2, a.shape[2], 0 if a.dtype == np.float16 else 1, a.shape[1], a.shape[0]))
            f.write(a.tobytes())
    if clut:
        a = rng.uniform(0.2, 0.45, (32, 192, 2)).astype(np.float16)          # six bands: four temperature anchors
        with open(os.path.join(MLV_DIR, "clut.lut"), "wb") as f:
            f.write(struct.pack("<IHBBII", 1234, 2, 2, 0, 192, 32))
            f.write(a.tobytes())
    return [ln % dict(dir=MLV_DIR) for ln in LUT_LINES + (CLUT_LINES if clut else [])]


def graph_goldens():
    """the module pass of the REFERENCE's own graph code over its own bin/default-darkroom.i-raw (oracle/ref_graph_shim.c: global.c,
    module.c, graph-io.c, connector.c, graph-export.c, graph-run-modules.h and the seven module main.c files compiled in place; only
    i-raw is a stand-in): every module on the path with image parameters, parameter block, connectors incl. negotiated sizes and
    request strengths, nodes, push constants, wiring and the committed uniform blocks.  pins the product's config reader, roi
    negotiation, module callbacks and commit_params (tests/test_host_ref_cpu.py)."""
    import gzip
    import json
    assert O.ref_host_lib() is not None, "oracle/_ref/libhostref.so missing: run `make -C oracle ref` where /root/reference exists"
    cases = [dict(lines=ln, w=w, h=h, raw=kw, text=O.ref_graph_describe(w, h, ln, kw)) for ln, w, h, kw in GRAPH_CASES]
    # bin/default-darkroom.i-mlv with the reference's own i-mlv/main.c reading a clip: header -> image parameters -> colour's block
    for name, (w, h, bpp, black, white, camera) in MLV_CLIPS.items():
        fn = write_golden_clip(name)
        lines = ["param:i-mlv:main:filename:" + fn] + (["param:denoise:01:strength:0.3"] if bpp == 12 else [])
        cases.append(dict(lines=lines, w=w, h=h, raw={}, mlv=name, text=O.ref_graph_describe(w, h, lines, {}, cfg="bin/default-darkroom.i-mlv")))
    # i-pfm (the reference's own i-pfm/main.c reading the header) in front of colour and filmcurv
    fn, cfg = write_golden_pfm()
    cases.append(dict(lines=[], w=64, h=40, raw={}, pfm=1, text=O.ref_graph_describe(64, 40, [], {}, cfg=cfg)))
    # the reference's own i-lut/main.c reading two tables into colour's abney / spectra connectors
    lines = write_golden_luts()
    cases.append(dict(lines=lines, w=640, h=480, raw={}, luts=1, text=O.ref_graph_describe(640, 480, lines, {})))
    lines = write_golden_luts(clut=True)    # with a clut the reference adds its autotemp node and a sink that feeds the gui
    cases.append(dict(lines=lines, w=640, h=480, raw={}, luts=2, text=O.ref_graph_describe(640, 480, lines, {})))
    with gzip.GzipFile(os.path.join(HERE, "host_graph.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(cases, indent=0).encode())
    print("graph goldens:", len(cases), "graphs,", sum(c["text"].count("\n") for c in cases), "lines")


CFG_LINES = [
    "param:colour:01:exposure:0.5", "param:crop:01:crop:0.1:0.9:0.2:0.8", "# comment", "param:nosuch:01:x:1", "param:colour:01:nosuch:1",
    "param:colour:09:exposure:1", "frames:7", "fps:30", "bogus:line", "module:grade:02", "module:nosuch:01", "connect:grade:01:output:grade:02:input",
    "connect:grade:01:nosuch:grade:02:input", "connect:nosuch:01:output:grade:02:input", "paramsub:colour:01:mat:4:0.5", "paramsub:colour:01:mat:40:0.5",
    "param:colour:01:mat:1:2:3", "param:colour:01:exposure", "param:denoise:01:strength:abc", "param:llap:01:sigma:0.3:0.4", "", "param", "param:",
    "connect:demosaic:01:output:colour:01:input", "connect:colour:01:output:colour:01:input", "connect:grade:01:output:crop:01:input",
    "param:i-raw:main:filename:some/long name with spaces.dng", "param:colour:01:import:abcdefghijklmnop", "paraminc:colour:01:exposure:0:0.25",
    "paramdec:colour:01:exposure:0:0.5", "module:grade:01", "param:grade:02:gain:2:2:2:2", "connect:-1:-1:-1:grade:02:input", "connect:hilite:01:output:grade:02:input",
    "param:colour:01:rbmap:0.1:0.2:0.3:0.4:0.5:0.6", "paramsub:colour:01:rbmap:140:1:2:3:4:5:6:7:8", "param:colour:01:cnt:7", "param:colour:01:cnt:7.9", "param:colour:01:matrix:-3",
    "module:llap:02:10:20", "connect:llap:02:output:llap:01:input", "connect:filmcurv:01:output:llap:02:input", "frames:-3", "frames:abc", "fps:", "param:crop:01:rotate:1e3",
    "param:crop:01:rotate:nan", "param:hilite:01:white:0x10", "param:filmcurv:01:light: 2.5", "module:display:dspy", "connect:llap:02:output:display:dspy:input",
]


def cfg_goldens():
    """the config grammar: return code of the REFERENCE's own dt_graph_read_config_line (graph-io.c compiled in place) for each of
    CFG_LINES on top of its bin/default-darkroom.i-raw, and what they did (frame count, all parameter blocks, all connections)."""
    import gzip
    import json
    codes, state = O.ref_config_lines(CFG_LINES)
    with gzip.GzipFile(os.path.join(HERE, "host_cfg.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(dict(lines=CFG_LINES, codes=codes, state=state), indent=0).encode())
    print("cfg goldens:", len(CFG_LINES), "lines, codes", sorted(set(codes)))


def shader_goldens():
    """outputs of the REFERENCE's own compute shaders, compiled as C++ (oracle/glsl, oracle/_ref/libshaderref.so), on the seeded inputs
    of tests/test_shader_ref_cpu.py: pins the oracle's float kernels where the reference is absent."""
    sys.path.insert(0, os.path.dirname(HERE))
    import test_shader_ref_cpu as T
    assert O.ref_shader_lib() is not None, "oracle/_ref/libshaderref.so missing: run `make -C oracle ref` where /root/reference exists"
    import json
    arrays, digests = {}, {}
    for name, fn in T.cases(O).items():
        _, got = fn()
        for k, g in enumerate(got):
            g = np.asarray(g, np.float32)
            if name.startswith(T.SAMPLED):      # compared with a one ulp allowance: the values are needed (all f16 valued)
                assert np.array_equal(g.astype(np.float16).astype(np.float32), g, equal_nan=True)
                arrays["%s/%d" % (name, k)] = g.astype(np.float16)
            else:                               # compared bit for bit: a digest is enough
                digests["%s/%d" % (name, k)] = T.digest(name, g)
    np.savez_compressed(os.path.join(HERE, "shader_ref.npz"), **arrays)
    with open(os.path.join(HERE, "shader_ref.json"), "w") as f:
        json.dump(digests, f, indent=0, sort_keys=True)
    print("shader goldens:", len(arrays), "images +", len(digests), "digests from", len(T.cases(O)), "cases")


def pipeline_goldens():
    """what the sink receives when the REFERENCE's own graph code and own shaders run on the CPU (oracle.ref_graph_describe +
    oracle.ref_pipeline_run): the default darkroom graph end to end, four configurations (tests/test_pipeline_ref_cpu.py)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import test_pipeline_ref_cpu as T
    out = {name: T.reference_output(O, name).astype(np.float32) for name in T.CASES}
    np.savez_compressed(os.path.join(HERE, "pipeline_ref.npz"), **out)
    for name in T.CASES:
        print("pipeline golden", name, "vs oracle: max abs %.3g, psnr %.1f dB" % T.check(name, out[name], T.oracle_output(O, name)))


def pipeline_goldens_2mp():
    """the same reference pipeline on the CPU at 3 MP (2004 x 1500: bayer + denoise, x-trans + denoise; ~1 min of shader
    emulation each).  the full output is 36 MB: the fixture keeps the lattice of every 8th pixel (offset 3) and the statistics of
    the whole frame against the oracle (tests/test_reference_gpu.py compares the product on the lattice)."""
    from oracle import oracle_py as O
    import test_pipeline_ref_cpu as T
    out = {}
    for name, (w, h, xtrans, strength) in T.CASES_2MP.items():
        raw = synth.mosaic(w, h, seed=77, xtrans=xtrans)
        lines = ["param:denoise:01:strength:%g" % strength]
        kw = dict(wb=T.WB, noise_a=T.NOISE[0], noise_b=T.NOISE[1])
        if xtrans:
            kw["filters"] = 9
        ref = O.ref_pipeline_run(O.ref_graph_describe(w, h, lines, kw), raw)[..., :3].astype(np.float32)
        d = O.darkroom_defaults(w, h)
        for k, v in enumerate(T.WB):
            d.whitebalance[k] = v
        d.noise_a, d.noise_b = T.NOISE
        d.denoise.strength = strength
        d.filters = 9 if xtrans else d.filters
        d.enable_grade = 1
        want = O.darkroom_run(d, raw)[..., :3]
        err = np.abs(ref.astype(np.float64) - want)
        mse = float(np.mean(err ** 2))
        el = err[3::8, 3::8]
        stats = np.array([ref.shape[0], ref.shape[1], err.max(), 99.0 if mse == 0 else 10 * np.log10(1.0 / mse), (err > 1e-3).mean(), (err != 0).mean(),
                          el.max(), (el > 1e-3).mean(), (el != 0).mean()])   # [6..8]: the same on the lattice
        print(name, ref.shape, "reference vs oracle over the whole frame: max abs %.3g psnr %.1f dB, > 1e-3: %.3g, differing: %.3g" % tuple(stats[2:6]))
        out[name] = ref[3::8, 3::8].copy()
        out[name + "_stats"] = stats
    np.savez_compressed(os.path.join(HERE, "pipeline_ref_2mp.npz"), **out)


def darkroom_goldens():
    w, h = 168, 126
    raw = synth.mosaic(w, h, seed=77)
    for name, strength in (("default", 0.0), ("denoise", 0.4)):
        d = O.darkroom_defaults(w, h)
        for k, v in enumerate((2.0, 1.0, 1.5)):
            d.whitebalance[k] = v
        d.denoise.strength = strength
        d.noise_a, d.noise_b = 100.0, 2.0
        out = O.darkroom_run(d, raw)
        np.savez_compressed(os.path.join(HERE, "darkroom_%s.npz" % name), raw=raw, out=out.astype(np.float32), strength=strength)
        print("darkroom golden", name, out.shape, float(out[..., :3].mean()))


if __name__ == "__main__":
    mlv_goldens()
    lj92_goldens()
    host_goldens()
    node_goldens()
    graph_goldens()
    cfg_goldens()
    shader_goldens()
    pipeline_goldens()
    if "--big" in sys.argv:
        pipeline_goldens_2mp()
    darkroom_goldens()
