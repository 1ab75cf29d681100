"""GPU parity tests, kernel by kernel and module by module: CUDA (through the C-ABI vkb_dispatch) vs the CPU oracle
on the same seeded inputs.  integer work is bit exact; float kernels are compared on their f16 outputs in ulps
(the oracle emulates the texture unit with exact float weights, the kernels use the analytic stencil weights and
SFU transcendentals: a few f16 roundings flip, nothing more)."""
import ctypes as C
import numpy as np
import pytest

from helpers import (assert_f16_close, assert_close_mixed, f16_ulp_diff, to_dev_f16, to_dev_u16, dev_f16, dev_f32, to_host, fbits, ibits,
                     ubits, psnr)
import plans
from vkdt_b200 import synth

pytestmark = pytest.mark.gpu


def _raw(w, h, seed=1, xtrans=False):
    return synth.mosaic(w, h, seed=seed, xtrans=xtrans)


def _noop_ref(O, raw, crop=(0, 0), ow=None, oh=None, black=2048.0, white=15000.0):
    h, w = raw.shape
    ow = ow or w; oh = oh or h
    src = (raw.astype(np.float32) / np.float32(65535.0))
    out, oi = O.new_img(oh, ow, 1)
    b = O.f4(*([np.float32(black) / np.float32(65535.0)] * 4)); wv = O.f4(*([np.float32(white) / np.float32(65535.0)] * 4))
    O.lib().o_denoise_noop(C.byref(O.img(src)), C.byref(oi), O.i4(crop[0], crop[1], 0, 0), b, wv)
    return out


@pytest.mark.parametrize("bpp", [10, 12, 14])
@pytest.mark.parametrize("n", [8, 2048, 2048 * 3 + 8, 4096 * 17 + 5, 1000003])
def test_mlv_unpack_bit_exact(gpu, oracle, bpp, n):
    rng = np.random.default_rng(bpp * 1000 + n % 997)
    pix = rng.integers(0, 1 << bpp, n, dtype=np.uint16)
    words = synth.pack_bits(pix, bpp)
    want = oracle.mlv_unpack(words, n, bpp)
    assert (want == pix).all()
    import torch
    pad = np.zeros((words.size + 7) // 8 * 8 + 8, dtype=np.uint16); pad[:words.size] = words
    d_in = to_dev_u16(pad)
    d_out = torch.zeros((n + 7) // 8 * 8, dtype=torch.int16, device="cuda")
    gpu.dispatch("i-mlv", "unpack", [gpu.image(d_in, pad.size, 1, 1, "ui16"), gpu.image(d_out, n, 1, 1, "ui16")], ibits(bpp))
    got = d_out.cpu().numpy().view(np.uint16)[:n]
    assert (got == pix).all()


@pytest.mark.parametrize("dims,crop", [((256, 192), (0, 0)), ((250, 190), (0, 0)), ((272, 200), (8, 4)), ((271, 201), (3, 5))])
def test_denoise_noop_bit_exact(gpu, oracle, dims, crop):
    w, h = dims
    raw = _raw(w, h)
    ow, oh = w - 2 * crop[0], h - 2 * crop[1]
    want = _noop_ref(oracle, raw, crop, ow, oh)
    d_in = to_dev_u16(raw); d_out = dev_f16(oh, ow)
    b, wh = np.float32(2048) / np.float32(65535), np.float32(15000) / np.float32(65535)
    push = ibits(crop[0], crop[1], crop[0] + ow, crop[1] + oh) + fbits(b, b, b, b, wh, wh, wh, wh, 0, 0, 0, 0) + ibits(0x5d5d5d5d, 0)
    gpu.dispatch("denoise", "noop", [gpu.image(d_in, w, h, 1, "ui16"), gpu.image(d_out, ow, oh, 1, "f16")], push, fbits(0, .6, 1, 0, 0, 0, 0, 0) + ibits(1))
    got = to_host(d_out)
    assert f16_ulp_diff(got, want).max() == 0


def test_rawnoop_fused_bit_exact(gpu, oracle):
    w, h = 512, 130
    raw = _raw(w, h)
    want = _noop_ref(oracle, raw)
    words = synth.pack_bits_fast14(raw)
    pad = np.zeros((words.size + 7) // 8 * 8 + 8, dtype=np.uint16); pad[:words.size] = words
    d_in = to_dev_u16(pad); d_out = dev_f16(h, w)
    b, wh = np.float32(2048) / np.float32(65535), np.float32(15000) / np.float32(65535)
    gpu.dispatch("b200", "rawnoop", [gpu.image(d_in, pad.size, 1, 1, "ui16"), gpu.image(d_out, w, h, 1, "f16")], ibits(14) + fbits(b, wh))
    assert f16_ulp_diff(to_host(d_out), want).max() == 0


def _mosaic_f16(oracle, w, h, seed=3, xtrans=False):
    return _noop_ref(oracle, _raw(w, h, seed, xtrans))


@pytest.mark.parametrize("dims", [(256, 192), (250, 186), (134, 102)])
def test_hilite_kernels(gpu, oracle, dims):
    O = oracle
    w, h = dims
    m = _mosaic_f16(O, w, h)
    hp = O.HiliteParams(0.985, 0.3, 0.6)
    par = fbits(0.985, 0.3, 0.6); push = fbits(1, 1, 1, 1) + ubits(0x5d5d5d5d)
    I = gpu.image
    # half
    want, wi = O.new_img(h // 2, w // 2, 4)
    O.lib().o_hilite_half(C.byref(O.img(m)), C.byref(wi), C.byref(hp), C.c_uint32(0x5d5d5d5d))
    d_m = to_dev_f16(m); d_half = dev_f16(h // 2, w // 2, 4)
    gpu.dispatch("hilite", "half", [I(d_m, w, h, 1, "f16"), I(d_half, w // 2, h // 2, 4, "f16")], push, par)
    assert f16_ulp_diff(to_host(d_half), want).max() == 0
    # reduce (fed with the oracle's half so that errors do not compound)
    rw, rh = (w // 2 - 1) // 2 + 1, (h // 2 - 1) // 2 + 1
    wred, wri = O.new_img(rh, rw, 4)
    O.lib().o_hilite_reduce(C.byref(wi), C.byref(wri), C.byref(hp), O.f4(1, 1, 1, 1))
    d_red = dev_f16(rh, rw, 4)
    gpu.dispatch("hilite", "reduce", [I(to_dev_f16(want), w // 2, h // 2, 4, "f16"), I(d_red, rw, rh, 4, "f16")], push, par)
    assert_f16_close(to_host(d_red), wred, 2, 0.97, "hilite reduce")
    # assemble
    wasm, wai = O.new_img(h // 2, w // 2, 4)
    O.lib().o_hilite_assemble(C.byref(wi), C.byref(wri), C.byref(wai), C.byref(hp))
    d_asm = dev_f16(h // 2, w // 2, 4)
    gpu.dispatch("hilite", "assemble", [I(to_dev_f16(want), w // 2, h // 2, 4, "f16"), I(to_dev_f16(wred), rw, rh, 4, "f16"),
                                        I(d_asm, w // 2, h // 2, 4, "f16")], push, par)
    assert_f16_close(to_host(d_asm), wasm, 2, 0.97, "hilite assemble")
    # doub
    wd, wdi = O.new_img(h, w, 1)
    O.lib().o_hilite_doub(C.byref(O.img(m)), C.byref(wai), C.byref(wdi), C.byref(hp), C.c_uint32(0x5d5d5d5d))
    d_d = dev_f16(h, w)
    gpu.dispatch("hilite", "doub", [I(d_m, w, h, 1, "f16"), I(to_dev_f16(wasm), w // 2, h // 2, 4, "f16"), I(d_d, w, h, 1, "f16")], push, par)
    assert_f16_close(to_host(d_d), wd, 1, 0.99, "hilite doub")


@pytest.mark.parametrize("dims", [(256, 192), (134, 102)])
def test_hilite_module(gpu, oracle, dims):
    O = oracle
    w, h = dims
    m = _mosaic_f16(O, w, h)
    hp = O.HiliteParams(0.985, 0.3, 0.6)
    want, wi = O.new_img(h, w, 1)
    O.lib().o_hilite_module(C.byref(O.img(m)), C.byref(wi), C.byref(hp), O.f4(1, 1, 1, 1), C.c_uint32(0x5d5d5d5d))
    got = to_host(plans.hilite(gpu, to_dev_f16(m), w, h, (0.985, 0.3, 0.6)))
    assert np.abs(got - want).max() < 2e-3 and psnr(got, want) > 70.0
    assert (m != want).mean() > 0.001  # the synthetic discs clip: hilite did reconstruct something


@pytest.mark.parametrize("dims", [(256, 192), (130, 98)])
def test_demosaic_kernels(gpu, oracle, dims):
    O = oracle
    w, h = dims
    m = _mosaic_f16(O, w, h)
    F = C.c_uint32(0x5d5d5d5d)
    cov, ci = O.new_img(h // 2, w // 2, 4); green, gi = O.new_img(h, w, 1); rgb, ri = O.new_img(h, w, 4)
    mi = O.img(m)
    O.lib().o_demosaic_gauss(C.byref(mi), C.byref(ci), F)
    O.lib().o_demosaic_splat(C.byref(mi), C.byref(ci), C.byref(gi), F)
    O.lib().o_demosaic_fix(C.byref(mi), C.byref(gi), C.byref(ci), C.byref(ri), F, 0)
    I = gpu.image
    push = fbits(1, 1, 1, 1) + ubits(0x5d5d5d5d)
    d_m = to_dev_f16(m)
    d_cov = dev_f16(h // 2, w // 2, 4)
    gpu.dispatch("demosaic", "gauss", [I(None, 0, 0, 1, "f16"), I(d_m, w, h, 1, "f16"), I(d_cov, w // 2, h // 2, 4, "f16")], push)
    gc = to_host(d_cov)
    # the eigenvector snap is a discontinuous decision: demand it agrees (almost) everywhere, eigenvalues in ulps
    same = (gc[..., 2:] == cov[..., 2:]).all(axis=-1)
    assert same.mean() > 0.999, "eigenvector snap agrees on %.5f of the blocks" % same.mean()
    assert_f16_close(gc[same][:, :2], cov[same][:, :2], 8, 0.90, "demosaic gauss eval")
    d_green = dev_f16(h, w)
    gpu.dispatch("demosaic", "splat", [I(d_m, w, h, 1, "f16"), I(to_dev_f16(cov), w // 2, h // 2, 4, "f16"), I(d_green, w, h, 1, "f16")], push)
    assert_f16_close(to_host(d_green), green, 2, 0.97, "demosaic splat")
    d_rgb = dev_f16(h, w, 4)
    gpu.dispatch("demosaic", "fix", [I(d_m, w, h, 1, "f16"), I(to_dev_f16(green), w, h, 1, "f16"), I(to_dev_f16(cov), w // 2, h // 2, 4, "f16"),
                                     I(d_rgb, w, h, 4, "f16")], push, ibits(0, 0))
    assert_f16_close(to_host(d_rgb), rgb, 2, 0.97, "demosaic fix")


def _rgb_image(w, h, seed=5, scale=1.0):
    img = synth.scene_rgb(w, h, seed) * np.float32(scale)
    a = np.ones((h, w, 4), dtype=np.float32); a[..., :3] = img
    return a.astype(np.float16).astype(np.float32)


def _colour_committed(O, d):
    f = np.zeros(242, dtype=np.float32)
    wb = (C.c_float * 4)(*d.colour.white)
    O.lib().o_colour_commit(C.byref(d.colour), wb, d.whitebalance, d.cam_to_rec2020, d.colour_primaries, d.colour_trc, O.fptr(f))
    return f


COLOUR_CASES = {
    "default": {},
    "wb_matrix_exposure": dict(wb=(2.0, 1.0, 1.5), mat=(0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8), exposure=0.7),
    "saturation": dict(wb=(1.8, 1.0, 1.4), sat=1.4),
    "rbf": dict(mode=1, cnt=4),
    "srgb_in_clip": dict(matrix=3, clip=1, clipmax=0.8),
}


@pytest.mark.parametrize("case", sorted(COLOUR_CASES))
def test_colour(gpu, oracle, case):
    O = oracle
    w, h = 192, 128
    cfg = COLOUR_CASES[case]
    d = O.darkroom_defaults(w, h)
    if "wb" in cfg:
        for k in range(3): d.whitebalance[k] = cfg["wb"][k]
    if "mat" in cfg:
        for k in range(9): d.cam_to_rec2020[k] = cfg["mat"][k]
    for k in ("exposure", "sat", "mode", "cnt", "matrix", "clip", "clipmax"):
        if k in cfg: setattr(d.colour, k, cfg[k])
    if case == "rbf":
        rb = [0.2, 0.3, 0.4, 0.25, 0.3, 0.35, 0.6, 0.5, 0.4, 0.55, 0.5, 0.45, 0.1, 0.8, 0.3, 0.12, 0.75, 0.3, 0.9, 0.9, 0.9, 0.95, 0.9, 0.85]
        for k, v in enumerate(rb): d.colour.rbmap[k] = v
    f = _colour_committed(O, d)
    a = _rgb_image(w, h)
    want, wi = O.new_img(h, w, 4)
    O.lib().o_colour_main(C.byref(O.img(a)), C.byref(wi), O.fptr(f), 1)
    d_out = dev_f16(h, w, 4)
    gpu.dispatch("colour", "main", [gpu.image(to_dev_f16(a), w, h, 4, "f16"), gpu.image(d_out, w, h, 4, "f16")], b"\0" * 12, f.tobytes())
    got = to_host(d_out)
    assert_f16_close(got[..., :3], want[..., :3], 3 if case == "saturation" else 2, 0.90, "colour " + case)


@pytest.mark.parametrize("trc,prim,clip", [(1, 1, 0), (3, 4, 0), (5, 6, 1), (7, 16, 0), (8, 14, 1), (9, 7, 0), (10, 8, 0), (11, 9, 1), (12, 10, 0),
                                           (13, 13, 0), (14, 11, 1), (14, 12, 0), (15, 15, 0)])
def test_colour_input_curves_and_gamuts(gpu, oracle, trc, prim, clip):
    """the input side of colour (main-impl.glsl:104-198): transfer curves 1..6, the camera log curves 7..15 of shared/oetf.glsl,
    the camera wide gamuts that go through xyz; the clip level runs through the same decode on the host."""
    O = oracle
    w, h = 192, 128
    d = O.darkroom_defaults(w, h)
    d.colour_primaries, d.colour_trc = prim, trc
    d.colour.clip, d.colour.clipmax = clip, 0.9
    f = _colour_committed(O, d)
    rng = np.random.default_rng(900 + 20 * trc + prim)
    a = rng.uniform(-0.05, 1.0, (h, w, 4)).astype(np.float16).astype(np.float32)
    a[1, :, :3] = np.linspace(-0.2, 1.2, w).astype(np.float16).astype(np.float32)[:, None]
    a[..., 3] = 1.0
    want, wi = O.new_img(h, w, 4)
    O.lib().o_colour_main(C.byref(O.img(a)), C.byref(wi), O.fptr(f), 1)
    d_out = dev_f16(h, w, 4)
    gpu.dispatch("colour", "main", [gpu.image(to_dev_f16(a), w, h, 4, "f16"), gpu.image(d_out, w, h, 4, "f16")], b"\0" * 12, f.tobytes())
    assert_f16_close(to_host(d_out)[..., :3], want[..., :3], 2, 0.90, "colour trc %d prim %d" % (trc, prim))


def _synthetic_luts(rng, nbands):
    """luts of the shapes the reference's offline tools write (clut: nbands squares of rg side by side; spectra: rgba, the
    dominant wavelength is -y / 2x; abney: rg with the gamut bounds in its last two columns)"""
    ch = 32
    h16 = lambda a: np.ascontiguousarray(a.astype(np.float16).astype(np.float32))
    clut = h16(rng.uniform(0.05, 0.6, (ch, nbands * ch, 2)))
    spectra = np.zeros((48, 48, 4), np.float32)
    sx = rng.uniform(0.5, 2.0, (48, 48)) * np.where(rng.uniform(0, 1, (48, 48)) < 0.5, -1.0, 1.0)
    lam = rng.uniform(380.0, 720.0, (48, 48))
    spectra[..., 0], spectra[..., 1], spectra[..., 2], spectra[..., 3] = sx, -2.0 * sx * lam, rng.uniform(0, 1, (48, 48)), rng.uniform(0.0, 0.9, (48, 48))
    abney = h16(rng.uniform(0.1, 0.6, (40, 64, 2)))
    abney[:, -2:, 0] = h16(rng.uniform(0.5, 0.9, (40, 2)))
    abney[:, -2:, 1] = h16(rng.uniform(0.6, 1.0, (40, 2)))
    return clut, np.ascontiguousarray(spectra), abney


@pytest.mark.parametrize("nbands,temp,use_clut,use_abney,sat,gamut,clip", [(3, 0.3, 1, 0, 1.0, 0, 0), (6, 0.62, 1, 0, 1.2, 0, 0), (3, 1.0, 1, 1, 1.3, 0, 0),
                                                                         (3, 0.0, 0, 1, 1.0, 1, 0), (3, 0.0, 0, 1, 1.4, 2, 0), (3, 0.0, 0, 1, 0.7, 3, 0),
                                                                         (6, 0.2, 1, 1, 1.5, 3, 1), (9, 0.85, 1, 0, 1.0, 0, 0)])
def test_colour_lut_inputs(gpu, oracle, nbands, temp, use_clut, use_abney, sat, gamut, clip):
    """colour with its lut connectors (main-impl.glsl:76-102 + clut.glsl: camera rgb -> rec2020 through a clut with temperature
    anchors; :287-335: saturation and gamut compression through the abney / spectra luts), the node's seven connectors and the
    push constants { have_clut, have_pick, have_abney } as colour/main.c:444-465 passes them."""
    import torch
    O = oracle
    w, h = 192, 128
    rng = np.random.default_rng(7000 + 10 * nbands + gamut)
    clut, spectra, abney = _synthetic_luts(rng, nbands)
    d = O.darkroom_defaults(w, h)
    d.colour.exposure, d.colour.sat, d.colour.matrix, d.colour.gamut, d.colour.clip, d.colour.clipmax = 0.2, sat, 4 if use_clut else 1, gamut, clip, 0.9
    f = _colour_committed(O, d)
    f[224] = temp
    a = rng.uniform(0.01, 1.2, (h, w, 4)).astype(np.float16).astype(np.float32)
    a[..., 3] = 1.0
    want, wi = O.new_img(h, w, 4)
    O.lib().o_colour_main_lut(C.byref(O.img(a)), C.byref(wi), O.fptr(f), 1, C.byref(O.img(clut)) if use_clut else None,
                              C.byref(O.img(abney)) if use_abney else None, C.byref(O.img(spectra)) if use_abney else None, C.c_float(0.0))
    d_in, d_out = to_dev_f16(a), dev_f16(h, w, 4)
    d_clut, d_abney, d_spec = to_dev_f16(clut), to_dev_f16(abney), torch.from_numpy(spectra).cuda()
    i_in = gpu.image(d_in, w, h, 4, "f16")
    conn = [i_in, gpu.image(d_out, w, h, 4, "f16"),
            gpu.image(d_clut, clut.shape[1], clut.shape[0], 2, "f16") if use_clut else i_in, i_in,
            gpu.image(d_abney, abney.shape[1], abney.shape[0], 2, "f16") if use_abney else i_in,
            gpu.image(d_spec, 48, 48, 4, "f32") if use_abney else i_in, i_in]
    gpu.dispatch("colour", "main", conn, np.array([use_clut, 0, use_abney], np.int32).tobytes(), f.tobytes())
    # the plain saturation runs through sin / cos / atan2 (not libm's bit for bit on the device): 3 ulps there, like test_colour
    assert_f16_close(to_host(d_out)[..., :3], want[..., :3], 3 if (sat != 1.0 and not use_abney) else 2, 0.90, "colour luts %d/%d/%d" % (nbands, use_clut, use_abney))


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_filmcurv(gpu, oracle, mode):
    O = oracle
    w, h = 192, 128
    a = _rgb_image(w, h, scale=1.3)
    fp = O.FilmcurvParams(3.0, 1.2, 0.01, mode, 1.2, 0.3, 0.1, -0.1, 0.2, 0.4)
    want, wi = O.new_img(h, w, 4)
    O.lib().o_filmcurv_main(C.byref(O.img(a)), C.byref(wi), C.byref(fp), 1)
    d_out = dev_f16(h, w, 4)
    gpu.dispatch("filmcurv", "main", [gpu.image(to_dev_f16(a), w, h, 4, "f16"), gpu.image(d_out, w, h, 4, "f16")], b"", bytes(fp))
    got = to_host(d_out)
    tol = {0: 6, 2: 6, 4: 6, 5: 6}.get(mode, 2)
    assert_close_mixed(got[..., :3], want[..., :3], tol, 4e-6, 0.85, "filmcurv mode %d" % mode)


@pytest.mark.parametrize("mode", [0, 1])
def test_grade(gpu, oracle, mode):
    O = oracle
    w, h = 160, 96
    a = _rgb_image(w, h)
    gp = O.GradeParams((C.c_float * 4)(0.01, 0.0, 0.02, 0.0), (C.c_float * 4)(1.1, 1.0, 0.9, 0.05), (C.c_float * 4)(1.0, 1.1, 0.95, 0.0),
                       (C.c_float * 4)(0.0, 0.01, 0.0, 0.0), mode, 0.3, 0.4)
    want, wi = O.new_img(h, w, 4)
    O.lib().o_grade_main(C.byref(O.img(a)), C.byref(wi), C.byref(gp), 0)
    d_out = dev_f32(h, w)
    gpu.dispatch("grade", "main", [gpu.image(to_dev_f16(a), w, h, 4, "f16"), gpu.image(d_out, w, h, 4, "f32")], b"", bytes(gp))
    got = to_host(d_out)
    assert np.abs(got[..., :3] - want[..., :3]).max() < 2e-5 * max(1.0, np.abs(want[..., :3]).max())


@pytest.mark.parametrize("dims,persp,rot", [((640, 480), None, 1337.0), ((320, 200), None, 1337.0), ((640, 480), (0.2, 0.22, 0.8, 0.25, 0.78, 0.8, 0.24, 0.76), 3.0)])
def test_crop_and_fused_chain(gpu, oracle, dims, persp, rot):
    """crop alone, then crop+colour+filmcurv as one kernel against the oracle's three node chain."""
    O = oracle
    w, h = dims
    d = O.darkroom_defaults(w, h)
    for k in range(3): d.whitebalance[k] = (2.0, 1.0, 1.5)[k]
    d.colour.exposure = 0.3
    d.crop.rotate = rot
    if persp:
        for k in range(8): d.crop.perspect[k] = persp[k]
        for k, v in enumerate((0.1, 0.9, 0.12, 0.88)): d.crop.crop[k] = v
    ow, oh = O.darkroom_out_size(d)
    fc = np.zeros(20, dtype=np.float32)
    O.lib().o_crop_commit(0, w, h, d.crop.perspect, d.crop.crop, C.byref(C.c_float(d.crop.rotate)), O.fptr(fc))
    a = _rgb_image(w, h)
    c0, c0i = O.new_img(oh, ow, 4); c1, c1i = O.new_img(oh, ow, 4); c2, c2i = O.new_img(oh, ow, 4)
    O.lib().o_crop_main(C.byref(O.img(a)), C.byref(c0i), O.fptr(fc))
    fcol = _colour_committed(O, d)
    O.lib().o_colour_main(C.byref(c0i), C.byref(c1i), O.fptr(fcol), 1)
    O.lib().o_filmcurv_main(C.byref(c1i), C.byref(c2i), C.byref(d.filmcurv), 1)
    d_a = to_dev_f16(a)
    d_out = dev_f16(oh, ow, 4)
    gpu.dispatch("crop", "main", [gpu.image(d_a, w, h, 4, "f16"), gpu.image(d_out, ow, oh, 4, "f16")], b"", fc.tobytes())
    got = to_host(d_out)
    if rot == 1337.0:
        assert f16_ulp_diff(got[..., :3], c0[..., :3]).max() == 0   # pure shifted copy
    else:
        assert_f16_close(got[..., :3], c0[..., :3], 2, 0.95, "crop catmull-rom")
    d_out2 = dev_f16(oh, ow, 4)
    gpu.dispatch("b200", "pointw", [gpu.image(d_a, w, h, 4, "f16"), gpu.image(d_out2, ow, oh, 4, "f16")], ubits(3, 1, 2, 3),
                 fc.tobytes() + fcol.tobytes() + bytes(d.filmcurv))
    assert_close_mixed(to_host(d_out2)[..., :3], c2[..., :3], 3, 4e-6, 0.90, "fused crop+colour+filmcurv", max_outliers=2e-5)
    if rot == 1337.0 and not persp:
        # the straight-line kernel the default parameters select against the node by node kernels with f16 images between
        # them (crop above, then colour, then filmcurv): the fused graph is value-compatible with the unfused one
        d_c = dev_f16(oh, ow, 4); d_f = dev_f16(oh, ow, 4)
        gpu.dispatch("colour", "main", [gpu.image(d_out, ow, oh, 4, "f16"), gpu.image(d_c, ow, oh, 4, "f16")], b"", fcol.tobytes())
        gpu.dispatch("filmcurv", "main", [gpu.image(d_c, ow, oh, 4, "f16"), gpu.image(d_f, ow, oh, 4, "f16")], b"", bytes(d.filmcurv))
        assert f16_ulp_diff(to_host(d_out2)[..., :3], to_host(d_f)[..., :3]).max() == 0


@pytest.mark.parametrize("dims", [(506, 384), (253, 191), (64, 40)])
@pytest.mark.parametrize("with_grade", [False, True])
def test_llap_module(gpu, oracle, dims, with_grade):
    O = oracle
    w, h = dims
    a = _rgb_image(w, h, scale=0.9)
    lp = O.LlapParams(0.12, 1.0, 1.0, 0.2)
    ll, lli = O.new_img(h, w, 4)
    O.lib().o_llap_module(C.byref(O.img(a)), C.byref(lli), C.byref(lp), 1)
    want = ll
    gbytes = None
    if with_grade:
        gp = O.GradeParams((C.c_float * 4)(0, 0, 0, 0), (C.c_float * 4)(1, 1, 1, 0), (C.c_float * 4)(1, 1, 1, 0), (C.c_float * 4)(0, 0, 0, 0), 0, 0.3, 0.4)
        want, wi = O.new_img(h, w, 4)
        O.lib().o_grade_main(C.byref(lli), C.byref(wi), C.byref(gp), 0)
        gbytes = bytes(gp)
    got = to_host(plans.llap(gpu, to_dev_f16(a), w, h, (0.12, 1.0, 1.0, 0.2), grade=gbytes, out_f32=with_grade))
    err = np.abs(got[..., :3] - want[..., :3])
    assert err.max() < 2e-3 and psnr(got[..., :3], want[..., :3]) > 66.0, (err.max(), psnr(got[..., :3], want[..., :3]))


def test_llapfin_against_the_exact_order_kernel(gpu):
    """(b200, llapfin) expands the coarse level with separable sums and runs two layers per packed instruction;
    (b200, llapfinx) is the per pixel kernel in the shader's 9-tap order.  they may differ by fp32 rounding before the
    f16 store, i.e. by one f16 ulp at a small fraction of the pixels, nowhere by more than two."""
    w, h = 506, 384
    d_in = to_dev_f16(_rgb_image(w, h, scale=0.9))
    a = to_host(plans.llap(gpu, d_in, w, h, (0.12, 1.0, 1.0, 0.2)))
    b = to_host(plans.llap(gpu, d_in, w, h, (0.12, 1.0, 1.0, 0.2), final_kernel="llapfinx"))
    ulp = f16_ulp_diff(a[..., :3], b[..., :3])
    assert ulp.max() <= 2 and (ulp > 0).mean() < 2e-3, (int(ulp.max()), float((ulp > 0).mean()))


# ---------------------------------------------------------------------------------------------------------
# wavelet denoise (denoise:strength > 0)
def _denoise_push(kind, wb, black, white, crop, filters, na, nb, level=0, block=2):
    b = [np.float32(black) / np.float32(65535.0)] * 4
    w = [np.float32(white) / np.float32(65535.0)] * 4
    head = fbits(*wb) + fbits(*b) + fbits(*w)
    if kind == "half":
        return head + ibits(*crop) + ubits(filters)
    if kind == "down":
        return head + ibits(0, 0, 0, 0) + fbits(na, nb) + ibits(level) + ubits(block)
    if kind == "assemble":
        return head + ibits(0, 0, 0, 0) + fbits(na, nb) + ubits(filters)
    if kind == "doub":
        return head + ibits(*crop) + ubits(filters) + fbits(na, nb) + ibits(0) + fbits(0, 0, 0, 0)
    raise ValueError(kind)


@pytest.mark.parametrize("dims,na,nb", [((256, 192), 100.0, 2.0), ((260, 196), 1e-5, 1e-4)])
def test_denoise_wavelet_kernels(gpu, oracle, dims, na, nb):
    """kernel by kernel, each fed with the oracle's previous stage; the second case uses tiny noise parameters so that the
    edge weights of down.comp:83-92 are not saturated."""
    O = oracle
    w, h = dims
    raw = _raw(w, h, seed=9)
    F = 0x5d5d5d5d
    wb = (2.0, 1.0, 1.5, 1.0)
    src = raw.astype(np.float32) / np.float32(65535.0)
    b4 = O.f4(*([np.float32(2048) / np.float32(65535)] * 4)); w4 = O.f4(*([np.float32(15000) / np.float32(65535)] * 4))
    dp = O.DenoiseParams(0.4, 0.6, 1.0, 0.0, (C.c_float * 4)(0, 0, 0, 0), 1)
    par = bytes(dp)
    hw, hh = w // 2, h // 2
    I = gpu.image
    d_raw = to_dev_u16(raw)
    # half
    half, hi_ = O.new_img(hh, hw, 4)
    O.lib().o_denoise_half(C.byref(O.img(src)), C.byref(hi_), O.i4(0, 0, w, h), w4, C.c_uint32(F))
    d_half = dev_f16(hh, hw, 4)
    gpu.dispatch("denoise", "half", [I(d_raw, w, h, 1, "ui16"), I(d_half, hw, hh, 4, "f16")], _denoise_push("half", wb, 2048, 15000, (0, 0, w, h), F, na, nb), par)
    assert f16_ulp_diff(to_host(d_half), half).max() == 0
    # downcov
    dn = [O.new_img(hh, hw, 4) for _ in range(4)]
    cov, ci = O.new_img(hh, hw, 4)
    O.lib().o_denoise_downcov(C.byref(hi_), C.byref(dn[0][1]), C.byref(ci))
    d_dn0 = dev_f16(hh, hw, 4); d_cov = dev_f16(hh, hw, 4)
    gpu.dispatch("denoise", "downcov", [I(to_dev_f16(half), hw, hh, 4, "f16"), I(d_dn0, hw, hh, 4, "f16"), I(d_cov, hw, hh, 4, "f16")],
                 _denoise_push("down", wb, 2048, 15000, None, F, na, nb, 0, 2), par)
    assert_close_mixed(to_host(d_cov), cov, 4, 1e-4, 0.90, "denoise downcov cov", max_outliers=1e-3, hard_atol=1.0)
    assert_close_mixed(to_host(d_dn0), dn[0][0], 3, 1e-5, 0.90, "denoise downcov", max_outliers=1e-3, hard_atol=5e-3)
    # down levels 1..3
    for lv in range(1, 4):
        O.lib().o_denoise_down(C.byref(dn[lv - 1][1]), C.byref(dn[lv][1]), C.byref(dp), b4, w4, C.c_float(na), C.c_float(nb), lv, C.c_uint32(2))
        d_out = dev_f16(hh, hw, 4)
        gpu.dispatch("denoise", "down", [I(to_dev_f16(dn[lv - 1][0]), hw, hh, 4, "f16"), I(d_out, hw, hh, 4, "f16")],
                     _denoise_push("down", wb, 2048, 15000, None, F, na, nb, lv, 2), par)
        assert_close_mixed(to_host(d_out), dn[lv][0], 2, 1e-6, 0.95, "denoise down level %d" % lv, max_outliers=1e-4)
    # assemble
    asm, ai = O.new_img(hh, hw, 4)
    O.lib().o_denoise_assemble(C.byref(hi_), C.byref(dn[0][1]), C.byref(dn[1][1]), C.byref(dn[2][1]), C.byref(dn[3][1]), C.byref(ai), C.byref(dp),
                               O.f4(*wb), b4, w4, C.c_float(na), C.c_float(nb), C.c_uint32(F))
    d_asm = dev_f16(hh, hw, 4)
    gpu.dispatch("denoise", "assemble", [I(to_dev_f16(half), hw, hh, 4, "f16")] + [I(to_dev_f16(dn[k][0]), hw, hh, 4, "f16") for k in range(4)] +
                 [I(d_asm, hw, hh, 4, "f16")], _denoise_push("assemble", wb, 2048, 15000, None, F, na, nb), par)
    assert_close_mixed(to_host(d_asm), asm, 3, 1e-5, 0.90, "denoise assemble", max_outliers=1e-4)
    # doub
    out, oi = O.new_img(h, w, 1)
    O.lib().o_denoise_doub(C.byref(O.img(src)), C.byref(ai), C.byref(hi_), C.byref(oi), C.byref(dp), O.i4(0, 0, w, h), b4, w4, C.c_float(na), C.c_float(nb), C.c_uint32(F))
    d_o = dev_f16(h, w)
    gpu.dispatch("denoise", "doub", [I(d_raw, w, h, 1, "ui16"), I(to_dev_f16(asm), hw, hh, 4, "f16"), I(to_dev_f16(half), hw, hh, 4, "f16"),
                                     I(d_o, w, h, 1, "f16")], _denoise_push("doub", wb, 2048, 15000, (0, 0, w, h), F, na, nb), par)
    assert_close_mixed(to_host(d_o), out, 2, 1e-6, 0.95, "denoise doub", max_outliers=1e-4)


# ---------------------------------------------------------------------------------------------------------
# x-trans (filters == 9): BASELINE config 3.  same kernels, block = 3 branches.
XT = 9


@pytest.mark.parametrize("dims", [(258, 192), (132, 102), (261, 195)])
def test_xtrans_hilite_and_demosaic_kernels(gpu, oracle, dims):
    O = oracle
    w, h = dims
    m = _mosaic_f16(O, w, h, seed=4, xtrans=True)
    hp = O.HiliteParams(0.985, 0.3, 0.6)
    par = fbits(0.985, 0.3, 0.6); push = fbits(1, 1, 1, 1) + ubits(XT)
    I = gpu.image
    mi = O.img(m)
    d_m = to_dev_f16(m)
    hw, hh = w // 3, h // 3
    # hilite half / doub on the 3x3 blocks
    half, hi_ = O.new_img(hh, hw, 4)
    O.lib().o_hilite_half(C.byref(mi), C.byref(hi_), C.byref(hp), C.c_uint32(XT))
    d_half = dev_f16(hh, hw, 4)
    gpu.dispatch("hilite", "half", [I(d_m, w, h, 1, "f16"), I(d_half, hw, hh, 4, "f16")], push, par)
    assert f16_ulp_diff(to_host(d_half), half).max() <= 1
    want, wi = O.new_img(h, w, 1)
    O.lib().o_hilite_module(C.byref(mi), C.byref(wi), C.byref(hp), O.f4(1, 1, 1, 1), C.c_uint32(XT))
    got = to_host(plans.hilite(gpu, d_m, w, h, (0.985, 0.3, 0.6), filters=XT))
    assert np.abs(got - want).max() < 2e-3 and psnr(got, want) > 70.0
    # demosaic
    cov, ci = O.new_img(hh, hw, 4); green, gi = O.new_img(h, w, 1); rgb, ri = O.new_img(h, w, 4)
    O.lib().o_demosaic_gauss(C.byref(mi), C.byref(ci), C.c_uint32(XT))
    O.lib().o_demosaic_splat(C.byref(mi), C.byref(ci), C.byref(gi), C.c_uint32(XT))
    O.lib().o_demosaic_fix(C.byref(mi), C.byref(gi), C.byref(ci), C.byref(ri), C.c_uint32(XT), 0)
    d_cov = dev_f16(hh, hw, 4)
    gpu.dispatch("demosaic", "gauss", [I(None, 0, 0, 1, "f16"), I(d_m, w, h, 1, "f16"), I(d_cov, hw, hh, 4, "f16")], push)
    gc = to_host(d_cov)
    same = (gc[..., 2:] == cov[..., 2:]).all(axis=-1)
    assert same.mean() > 0.999
    d_green = dev_f16(h, w)
    gpu.dispatch("demosaic", "splat", [I(d_m, w, h, 1, "f16"), I(to_dev_f16(cov), hw, hh, 4, "f16"), I(d_green, w, h, 1, "f16")], push)
    assert_f16_close(to_host(d_green), green, 2, 0.97, "xtrans splat")
    d_rgb = dev_f16(h, w, 4)
    gpu.dispatch("demosaic", "fix", [I(d_m, w, h, 1, "f16"), I(to_dev_f16(green), w, h, 1, "f16"), I(to_dev_f16(cov), hw, hh, 4, "f16"),
                                     I(d_rgb, w, h, 4, "f16")], push, ibits(0, 0))
    assert_f16_close(to_host(d_rgb), rgb, 2, 0.97, "xtrans fix")
