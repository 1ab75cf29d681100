"""lossless jpeg (LJ92) decoder behind lossless MLV clips (video_mlv.c:224-251).  Host code, no GPU.
tests/golden/lj92.npz holds streams written by the reference's own liblj92 encoder and by our test encoder, each with the
output of the REFERENCE's decoder, and a lossless clip with the frames the reference's mlv_get_frame returned."""
import os
import numpy as np
import pytest

from vkdt_b200 import api, synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "lj92.npz"))


@pytest.mark.parametrize("i", range(10))
def test_decoder_matches_reference_liblj92(i):
    out, bits = api.lj92_decode(G["stream_%d" % i].tobytes())
    want = G["expected_%d" % i]
    assert out.shape == want.shape and np.array_equal(out, want), str(G["notes"][i])


@pytest.mark.parametrize("bits,comps,pred", [(16, 1, 7), (14, 2, 6), (14, 4, 4), (12, 2, 5), (10, 1, 2), (14, 1, 3), (2, 1, 1)])
def test_roundtrip_t81(bits, comps, pred):
    """T.81 annex H semantics for what the goldens cannot cover (16 bit overflows the vendored decoder, its multi component
    row loops only implement predictor 1): encode -> decode is the identity, including the 32768 difference class."""
    rng = np.random.default_rng(bits * 100 + comps * 10 + pred)
    img = rng.integers(0, 1 << bits, (17, 12 * comps)).astype(np.uint16)
    if bits == 16:
        img[5, 3], img[5, 4] = 0, 32768     # difference of exactly 32768: category 16, no extra bits
    out, b = api.lj92_decode(synth.lj92_encode(img, bits, comps, pred))
    assert b == bits and np.array_equal(out, img)


def test_rejects_garbage():
    s = bytearray(synth.lj92_encode(np.arange(64, dtype=np.uint16).reshape(8, 8), 14))
    with pytest.raises(api.VkbError):
        api.lj92_decode(bytes(s[:40]))                 # truncated before the scan
    with pytest.raises(api.VkbError):
        api.lj92_decode(b"\xff\xd8\xff\xc0\x00\x04\x00\x00")   # baseline dct frame: another process
    s[3] = 0xC4 ^ 0x01
    with pytest.raises(api.VkbError):
        api.lj92_decode(bytes(s))


def test_lossless_clip_plan(tmp_path):
    fn = tmp_path / "l.mlv"
    fn.write_bytes(G["clip"].tobytes())
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-mlv"))
    assert g.line("param:i-mlv:main:filename:%s" % fn) == 0
    text = g.plan()
    # decoded on the host, uploaded as plain u16: no device unpack
    assert "source i-mlv bytes %d packed 0" % (64 * 34 * 2) in text and "unpack" not in text and "rawnoop" not in text
