"""CPU tests of the oracle (test infrastructure): pinned against the reference-made MLV goldens, against the
reference's own decoder when oracle/_ref was built, and against its own committed darkroom outputs."""
import ctypes as C
import os
import numpy as np
import pytest

from vkdt_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("bpp", [10, 12, 14])
def test_mlv_unpack_matches_reference_golden(oracle, bpp):
    g = np.load(os.path.join(GOLD, "mlv_unpack_%d.npz" % bpp))
    got = oracle.mlv_unpack(g["words"], int(g["width"]) * int(g["height"]), bpp)
    assert (got.reshape(g["expected"].shape) == g["expected"]).all()


@pytest.mark.parametrize("bpp", [10, 12, 14])
def test_mlv_unpack_matches_live_reference(oracle, bpp, tmp_path):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    rng = np.random.default_rng(bpp)
    pix = rng.integers(0, 1 << bpp, (38, 72), dtype=np.uint16)
    fn = str(tmp_path / "t.mlv")
    synth.write_mlv(fn, [pix, pix[::-1].copy()], bpp=bpp)
    for f, want in ((0, pix), (1, pix[::-1])):
        out, info = oracle.ref_mlv_decode(fn, f)
        assert info["frames"] == 2 and (out == want).all()
        assert (oracle.mlv_unpack(synth.pack_bits(want, bpp), want.size, bpp).reshape(want.shape) == out).all()


def test_pack_roundtrip_ragged():
    for n in (1, 7, 9, 1001):
        for bpp in (10, 12, 14):
            pix = np.random.default_rng(n).integers(0, 1 << bpp, n, dtype=np.uint16)
            from oracle import oracle_py as O
            assert (O.mlv_unpack(synth.pack_bits(pix, bpp), n, bpp) == pix).all()


def test_f16_rounding_matches_numpy(oracle):
    lib = oracle.lib()
    lib.o_darkroom_defaults  # loaded
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1.0, 300.0, 7e4)] +
                       [np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e6, 2.0 ** -24, 2.0 ** -25, 1.5 * 2.0 ** -25, np.inf], dtype=np.float32)])
    # route through the oracle's store: denoise noop with black 0 white 1 on unorm input is the identity + f16 store
    src = np.abs(x).reshape(1, -1)
    out, oi = oracle.new_img(1, src.shape[1], 1)
    lib.o_denoise_noop(C.byref(oracle.img(src)), C.byref(oi), oracle.i4(0, 0, 0, 0), oracle.f4(0, 0, 0, 0), oracle.f4(1, 1, 1, 1))
    with np.errstate(over="ignore"):
        want = src.astype(np.float16).astype(np.float32)
    assert np.array_equal(out, want)


@pytest.mark.parametrize("name", ["default", "denoise"])
def test_darkroom_matches_committed_golden(oracle, name):
    g = np.load(os.path.join(GOLD, "darkroom_%s.npz" % name))
    raw = g["raw"]
    d = oracle.darkroom_defaults(raw.shape[1], raw.shape[0])
    for k, v in enumerate((2.0, 1.0, 1.5)):
        d.whitebalance[k] = v
    d.denoise.strength = float(g["strength"])
    d.noise_a, d.noise_b = 100.0, 2.0
    out = oracle.darkroom_run(d, raw)
    assert out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() <= 1e-6


def test_crop_micro_crop_sizes(oracle):
    """crop/main.c:194-224,270-274: 3 px per side when the dimension exceeds 400, sizes fall out of float math."""
    for (w, h), want in (((512, 384), (506, 384)), ((6000, 4000), None), ((4096, 2160), (4090, 2154)), ((300, 200), (300, 200))):
        d = oracle.darkroom_defaults(w, h)
        ow, oh = oracle.darkroom_out_size(d)
        assert w - 7 <= ow <= w and h - 7 <= oh <= h
        if want:
            assert (ow, oh) == want


def test_pfm_writer(oracle, tmp_path):
    rgba = np.arange(5 * 3 * 4, dtype=np.float32).reshape(3, 5, 4)
    fn = str(tmp_path / "x.pfm")
    assert oracle.lib().o_write_pfm(fn.encode(), oracle.fptr(rgba), 5, 3) == 0
    data = open(fn, "rb").read()
    assert data.startswith(b"PF\n5 3\n-1.0")
    hdr_end = data.index(b"\n", data.index(b"-1.0")) + 1
    assert hdr_end % 16 == 0                       # o-pfm/main.c:27-34: payload 16-byte aligned
    px = np.frombuffer(data[hdr_end:], dtype=np.float32).reshape(3, 5, 3)
    assert np.array_equal(px, rgba[..., :3])       # rgb only, no y flip


def test_sampler_mirrored_repeat(oracle):
    """texture() of the oracle: mirrored repeat, bilinear; reduce of a constant image is the constant."""
    a = np.full((9, 13), 0.25, dtype=np.float32)
    out, oi = oracle.new_img(5, 7, 1)
    oracle.lib().o_llap_reduce(C.byref(oracle.img(a)), C.byref(oi))
    assert np.array_equal(out, np.full((5, 7), 0.25, dtype=np.float32))
