"""i-raw file path: uncompressed cfa dng -> image parameters.  Host logic only, no GPU.
Follows i-raw/rawloader-c/lib.rs:137-279 (window alignment, crop box) and i-raw/main.c:138-256 (wb, colour matrix)."""
import numpy as np
import pytest

from vkdt_b200 import api, synth

XYZ_TO_REC2020 = np.array([[1.71665119, -0.35567078, -0.25336628], [-0.66668435, 1.61648124, 0.01576855],
                           [0.01763986, -0.04277061, 0.94210312]])
CM = (0.9, -0.3, -0.1, -0.4, 1.2, 0.2, -0.1, 0.2, 0.7)


def _raw(w, h, seed=1):
    return np.random.default_rng(seed).integers(0, 65535, (h, w), dtype=np.uint16)


@pytest.mark.parametrize("shift", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("big_endian", [False, True])
def test_bayer_alignment(tmp_path, shift, big_endian):
    sy, sx = shift
    w, h = 37, 23
    stored = np.roll(np.array([[0, 1], [1, 2]]), (-sy, -sx), axis=(0, 1))  # pattern as the file stores it
    fn = str(tmp_path / "a.dng")
    synth.write_dng(fn, _raw(w, h), cfa=stored, black=(100, 200, 300, 400), white=16000, neutral=(0.5, 1.0, 0.8),
                    color_matrix=CM, big_endian=big_endian)
    p, ox, oy = api.dng_info(fn)
    # the emitted window starts on red
    assert stored[oy % 2, ox % 2] == 0 and stored[oy % 2, (ox + 1) % 2] == 1
    assert (ox, oy) == (sx, sy)
    assert p.width == ((w - ox) // 2) * 2 and p.height == ((h - oy) // 2) * 2
    assert p.filters not in (0, 9)
    assert list(p.crop_aabb) == [0, 0, ((w - ox) // 2) * 2, ((h - oy) // 2) * 2]
    # black levels follow the window: index k of the aligned 2x2 block reads the stored block at the shifted phase
    bl = np.array([[100, 200], [300, 400]], np.float32)
    assert list(p.black) == [bl[(r + oy) % 2, (c + ox) % 2] for r in (0, 1) for c in (0, 1)]
    assert list(p.white) == [16000.0] * 4
    np.testing.assert_allclose(list(p.whitebalance), [2.0, 1.0, 1.25, 1.0], rtol=1e-6)
    want = XYZ_TO_REC2020 @ np.linalg.inv(np.array(CM).reshape(3, 3))
    np.testing.assert_allclose(np.array(list(p.cam_to_rec2020)).reshape(3, 3), want, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("sy", range(6))
@pytest.mark.parametrize("sx", range(6))
def test_xtrans_alignment(tmp_path, sx, sy):
    w, h = 50, 44
    stored = np.roll(synth.XTRANS, (-sy, -sx), axis=(0, 1))
    fn = str(tmp_path / "x.dng")
    synth.write_dng(fn, _raw(w, h), cfa=stored)
    p, ox, oy = api.dng_info(fn)
    assert p.filters == 9
    # whatever the stored phase, the window begins on the canonical layout the kernels assume
    aligned = np.roll(stored, (-oy, -ox), axis=(0, 1))
    assert (aligned == synth.XTRANS).all(), (ox, oy)
    assert p.width == ((w - ox) // 3) * 3 and p.height == ((h - oy) // 3) * 3
    assert p.width + ox <= w and p.height + oy <= h


def test_active_area_and_illuminant(tmp_path):
    w, h = 64, 48
    fn = str(tmp_path / "c.dng")
    # gbrg file: window moves down by one row; active area (top left bottom right)
    synth.write_dng(fn, _raw(w, h), cfa=((1, 2), (0, 1)), active_area=(3, 5, 45, 61), illuminant=17, orientation=6,
                    color_matrix=CM)
    p, ox, oy = api.dng_info(fn)
    assert (ox, oy) == (0, 1)
    # lib.rs:263-270: x y round up to the pattern, X Y move with the window and round down to the block
    assert list(p.crop_aabb) == [6, 4, 60, 44]
    assert p.orientation == 6
    a_to_d65 = np.array([[9.50674182e-01, -1.87430902e-01, 2.62831155e-01], [-2.56724729e-02, 1.03231456e+00, -1.15608371e-02],
                         [-2.74089665e-03, 9.09809774e-02, 2.81290019e+00]])
    want = XYZ_TO_REC2020 @ np.linalg.inv(np.array(CM).reshape(3, 3) @ a_to_d65)
    np.testing.assert_allclose(np.array(list(p.cam_to_rec2020)).reshape(3, 3), want, rtol=5e-4, atol=5e-5)


def test_rejects_what_it_cannot_decode(tmp_path):
    fn = str(tmp_path / "n.dng")
    open(fn, "wb").write(b"II*\0" + b"\0" * 64)
    with pytest.raises(api.VkbError):
        api.dng_info(fn)
    with pytest.raises(api.VkbError):
        api.dng_info(str(tmp_path / "missing.dng"))


def test_graph_resolves_dng_source(tmp_path):
    w, h = 403, 301
    fn = str(tmp_path / "g.dng")
    synth.write_dng(fn, _raw(w, h), cfa=((2, 1), (1, 0)))   # bggr: window starts at (1, 1)
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    assert g.line("param:i-raw:main:filename:%s" % fn) == 0
    text = g.plan()
    assert "source i-raw" in text
    ow, oh = g.sink_size("main")
    assert 0 < ow <= 402 and 0 < oh <= 300
    g2 = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    assert g2.line("param:i-raw:main:filename:%s" % str(tmp_path / "g.cr2")) == 0
    with pytest.raises(api.VkbError):
        g2.plan()
