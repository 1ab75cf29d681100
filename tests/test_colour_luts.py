"""colour's lut inputs end to end: `i-lut` modules (i-lut/main.c, core/lut.h) wired to colour's clut / abney /
spectra connectors (colour/main.c:416-465), the temperature anchor blend of colour/main.c:268-292 or the autotemp node's as-shot
answer (atemp-impl.glsl), the (colour, main) launch with its seven connectors.  the per pixel arithmetic is pinned against the reference's shader in tests/test_shader_ref_cpu.py; here
the planner's wiring on the host and the developed frame against the oracle on the GPU.  the tables are synthetic (the
reference's own are made by its offline tools from measured camera data and do not come with a checkout)."""
import ctypes as C
import os
import struct
import numpy as np
import pytest

from vkdt_b200 import api, synth
from helpers import parity_gate, psnr

WB, CAM = (2.0, 1.0, 1.5), (0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8)


def synthetic_luts(rng, nbands):
    ch = 32
    clut = rng.uniform(0.2, 0.45, (ch, nbands * ch, 2)).astype(np.float16)
    spectra = np.zeros((48, 48, 4), np.float32)
    sx = rng.uniform(0.5, 2.0, (48, 48)) * np.where(rng.uniform(0, 1, (48, 48)) < 0.5, -1.0, 1.0)
    lam = rng.uniform(380.0, 720.0, (48, 48))
    spectra[..., 0], spectra[..., 1], spectra[..., 2], spectra[..., 3] = sx, -2.0 * sx * lam, rng.uniform(0, 1, (48, 48)), rng.uniform(0.0, 0.9, (48, 48))
    abney = rng.uniform(0.25, 0.4, (40, 64, 2)).astype(np.float16)
    abney[:, -2:, 0] = rng.uniform(0.5, 0.9, (40, 2)).astype(np.float16)
    abney[:, -2:, 1] = rng.uniform(0.6, 1.0, (40, 2)).astype(np.float16)
    return clut, np.ascontiguousarray(spectra), abney


def write_lut(path, a, version=2, magic=1234):
    """core/lut.h: { u32 magic, u16 version, u8 channels, u8 datatype (0 half, 1 float), u32 wd, u32 ht } + texels"""
    h, w, c = a.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<IHBBII", magic, version, c, 0 if a.dtype == np.float16 else 1, w, h))
        f.write(np.ascontiguousarray(a).tobytes())


def lut_lines(tmp, which):
    out = []
    for name in which:
        out += ["module:i-lut:%s" % name, "param:i-lut:%s:filename:%s" % (name, os.path.join(tmp, name + ".lut")),
                "connect:i-lut:%s:output:colour:01:%s" % (name, name)]
    return out


def test_planner_wires_the_luts_as_sources(tmp_path):
    clut, spectra, abney = synthetic_luts(np.random.default_rng(1), 6)
    write_lut(tmp_path / "clut.lut", clut); write_lut(tmp_path / "abney.lut", abney); write_lut(tmp_path / "spectra.lut", spectra)
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    for ln in lut_lines(str(tmp_path), ("clut", "abney", "spectra")) + ["param:colour:01:matrix:4", "param:colour:01:temp:5000"]:
        assert g.line(ln) == 0, ln
    raw = np.zeros((384, 512), dtype=np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(512, 384))
    plan = g.plan()
    col = [ln for ln in plan.splitlines() if "colour_main" in ln]
    assert len(col) == 1, plan                     # a launch of its own: the fused pointwise chain does not read luts
    assert plan.count("launch 21 colour_autotemp ") == 1 and ":1x1x1x1:f32" in col[0]   # the as-shot temperature node comes with a clut
    assert ":192x32x2x1:f16" in col[0] and ":64x40x2x1:f16" in col[0] and ":48x48x4x1:f32" in col[0], col[0]
    assert plan.count("source i-lut") == 3 and "source i-lut bytes %d" % spectra.nbytes in plan
    g.close()
    # without luts the chain stays fused
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    g.set_source(raw.ctypes.data, api.raw_params(512, 384))
    assert "colour_main" not in g.plan() and "b200_pointw" in g.plan()
    g.close()


@pytest.mark.parametrize("what", ["missing", "magic", "version", "short", "truncated", "huge"])
def test_a_lut_that_cannot_be_read_fails_the_plan(tmp_path, what):
    clut, spectra, abney = synthetic_luts(np.random.default_rng(2), 3)
    write_lut(tmp_path / "abney.lut", abney)
    if what == "magic": write_lut(tmp_path / "spectra.lut", spectra, magic=4321)
    elif what == "version": write_lut(tmp_path / "spectra.lut", spectra, version=1)
    elif what == "short": open(tmp_path / "spectra.lut", "wb").write(b"\xd2\x04\0\0\x02\0")
    elif what == "truncated":
        write_lut(tmp_path / "spectra.lut", spectra)
        blob = open(tmp_path / "spectra.lut", "rb").read()
        open(tmp_path / "spectra.lut", "wb").write(blob[:len(blob) // 2])
    elif what == "huge": open(tmp_path / "spectra.lut", "wb").write(struct.pack("<IHBBII", 1234, 2, 4, 1, 60000, 60000) + b"\0" * 64)
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    for ln in lut_lines(str(tmp_path), ("abney", "spectra")):
        assert g.line(ln) == 0, ln
    raw = np.zeros((384, 512), dtype=np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(512, 384))
    with pytest.raises(api.VkbError):
        g.plan()
    g.close()


CASES = {
    "abney-rec2020":  (("abney", "spectra"), 3, ["param:colour:01:gamut:2", "param:colour:01:sat:1.3"], [("colour.gamut", 2), ("colour.sat", 1.3)]),
    "abney-locus":    (("abney", "spectra"), 3, ["param:colour:01:gamut:1"], [("colour.gamut", 1)]),
    "clut-legacy":    (("clut",), 3, ["param:colour:01:matrix:4", "param:colour:01:temp:4500"], [("colour.matrix", 4), ("colour.temp", 4500.0)]),
    "clut-as-shot":   (("clut",), 6, ["param:colour:01:matrix:4", "param:colour:01:temp:0"], [("colour.matrix", 4), ("colour.temp", 0.0)]),
    "clut-anchors":   (("clut", "abney", "spectra"), 6, ["param:colour:01:matrix:4", "param:colour:01:temp:5200", "param:colour:01:gamut:3", "param:colour:01:sat:0.8"],
                       [("colour.matrix", 4), ("colour.temp", 5200.0), ("colour.gamut", 3), ("colour.sat", 0.8)]),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_luts_end_to_end(gpu, oracle, tmp_path, name):
    from test_graph_gpu import _oracle_cfg, _run_graph, _set
    which, nbands, lines, fields = CASES[name]
    w, h = 512, 384
    clut, spectra, abney = synthetic_luts(np.random.default_rng(5), nbands)
    write_lut(tmp_path / "clut.lut", clut); write_lut(tmp_path / "abney.lut", abney); write_lut(tmp_path / "spectra.lut", spectra)
    raw = synth.mosaic(w, h, seed=41)
    d = _oracle_cfg(oracle, w, h)
    for path, val in fields:
        _set(d, path, val)
    O = oracle
    imgs = {"clut": O.img(clut.astype(np.float32)), "abney": O.img(abney.astype(np.float32)), "spectra": O.img(spectra)}
    O.lib().o_set_colour_luts(C.byref(imgs["clut"]) if "clut" in which else None, C.byref(imgs["abney"]) if "abney" in which else None,
                              C.byref(imgs["spectra"]) if "spectra" in which else None)
    try:
        want = oracle.darkroom_run(d, raw)
    finally:
        O.lib().o_set_colour_luts(None, None, None)
    got, g = _run_graph(gpu, raw, extra=tuple(lut_lines(str(tmp_path), which) + lines))
    assert "colour_main" in g.plan()
    assert got.shape == want.shape and np.isfinite(got[..., :3]).all()
    err = np.abs(got[..., :3] - want[..., :3])
    print("%s: max abs %.3g psnr %.1f" % (name, err.max(), psnr(got[..., :3], want[..., :3])))
    parity_gate(got[..., :3], want[..., :3], "luts " + name)
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nbands", [3, 6, 9])
def test_autotemp_kernel(gpu, oracle, nbands):
    """(colour, autotemp) on its own: clut, 1x1 f32 answer, picked (a dummy) against the restatement of atemp-impl.glsl"""
    import torch
    from helpers import to_dev_f16
    O = oracle
    clut, _s, _a = synthetic_luts(np.random.default_rng(8 + nbands), nbands)
    clut = clut.astype(np.float32)
    for b in range(nbands):
        clut[:, b * 32:(b + 1) * 32, 0] += np.float32(0.04 * b - 0.1)
    clut = np.ascontiguousarray(clut.astype(np.float16).astype(np.float32))
    O.lib().o_colour_autotemp.restype = C.c_float
    for temp, wb in ((0.0, (2.1, 1.0, 1.6)), (0.0, (1.3, 1.0, 2.4)), (5000.0, (2.1, 1.0, 1.6))):
        d = O.darkroom_defaults(64, 64)
        d.colour.temp, d.colour.matrix = temp, 4
        for k in range(3): d.whitebalance[k] = wb[k]
        f = np.zeros(242, dtype=np.float32)
        O.lib().o_colour_commit(C.byref(d.colour), (C.c_float * 4)(*d.colour.white), d.whitebalance, d.cam_to_rec2020, 0, 0, O.fptr(f))
        want = O.lib().o_colour_autotemp(C.byref(O.img(clut)), O.fptr(f))
        d_clut, d_out = to_dev_f16(clut), torch.zeros(1, dtype=torch.float32, device="cuda")
        i_clut = gpu.image(d_clut, clut.shape[1], clut.shape[0], 2, "f16")
        gpu.dispatch("colour", "autotemp", [i_clut, gpu.image(d_out, 1, 1, 1, "f32"), i_clut], np.zeros(1, np.int32).tobytes(), f.tobytes())
        got = float(d_out.cpu()[0])
        assert got == want, (nbands, temp, wb, got, want)
        assert (temp > 0) == (want == -1.0)
