"""DNG GainMap opcodes: OpcodeList2 of a dng -> four Bayer gain maps -> a (denoise, gainmap) source node
-> noop.comp:48-57 / doub.comp:106-114.  the opcode list decoder against the reference's own i-raw/dng_opcode_decode.c compiled
in place (oracle/_ref/libdngopref.so; tests/golden/dngop.json keeps its answers), the planner's wiring on the host, and the
developed frame against the oracle on the GPU, from a dng file and from memory."""
import ctypes as C
import json
import os
import struct
import numpy as np
import pytest

from vkdt_b200 import api, synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "dngop.json")
WB, CAM = (2.0, 1.0, 1.5), (0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8)


def gain_maps(h, w, pv=7, ph=9, seed=3, top=0, left=0):
    """four GainMap opcodes, one per site of the 2x2 block (rggb order), smooth vignetting-like gains in [1, 1.6]"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:pv, 0:ph]
    r2 = ((xx / (ph - 1) - 0.5) ** 2 + (yy / (pv - 1) - 0.5) ** 2)
    planes, ops = [], []
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        g = (1.0 + (0.8 + 0.1 * k) * r2 + rng.uniform(0, 0.02, (pv, ph))).astype(np.float32)
        planes.append(g)
        ops.append(synth.dng_gain_map_opcode(g, top + dy, left + dx, h, w))
    return planes, synth.dng_opcode_list(ops)


def blobs():
    """opcode lists the decoders have to agree on: gain maps, foreign opcodes in between, truncated and inconsistent lists"""
    _, good = gain_maps(64, 96)
    other = struct.pack(">4I", 6, 0x01030000, 1, 16) + struct.pack(">4I", 2, 4, 60, 92)          # TrimBounds
    unknown = struct.pack(">4I", 77, 0x01040000, 3, 8) + b"\x01\x02\x03\x04\x05\x06\x07\x08"
    g1 = synth.dng_gain_map_opcode(np.ones((3, 2), np.float32) * 1.25, 2, 4, 64, 96, spacing=(0.5, 1.0), origin=(0.125, 0.25))
    bad_len = bytearray(g1); bad_len[15] += 4          # declared size 4 bytes too long
    return {
        "four gain maps": good,
        "mixed": synth.dng_opcode_list([other, g1, unknown]),
        "empty": struct.pack(">I", 0),
        "truncated": good[:-5],
        "trailing": good + b"\0\0",
        "gain map with a wrong size": synth.dng_opcode_list([bytes(bad_len) + b"\0\0\0\0"]),
        "short": b"\0\0",
    }


def _product(blob):
    api.lib.vkb_dng_opcodes_describe.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    buf = C.create_string_buffer(1 << 18)
    api.check(api.lib.vkb_dng_opcodes_describe(bytes(blob), len(blob), buf, len(buf)))
    return buf.value.decode()


def _reference(blob):
    so = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libdngopref.so")
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    buf = C.create_string_buffer(1 << 18)
    b = bytes(blob) + b"\0" * 16        # the reference reads a list's count before it looks at the length
    L.ref_dngop_describe(b, len(blob), buf, len(buf))
    return buf.value.decode()


def test_opcode_list_decoder_matches_the_reference_decoder():
    golden = json.load(open(GOLDEN))
    for name, blob in blobs().items():
        got = _product(blob)
        assert got == golden[name], name
        if name not in ("short",):          # (the reference reads four bytes whatever the length)
            ref = _reference(blob)
            if ref is not None:
                assert got == ref, name


def _cfg(strength):
    return api.DARKROOM_CFG.format(src="i-raw") + "param:denoise:01:strength:%g\n" % strength


@pytest.mark.parametrize("strength", [0.0, 0.4])
def test_gain_maps_become_a_source_node_of_denoise(strength):
    w, h = 128, 96
    _, blob = gain_maps(h, w)
    g = api.Graph(cfg_text=_cfg(strength))
    raw = np.zeros((h, w), np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(w, h))
    g.set_dng_opcodes(blob)
    plan = g.plan()
    assert ("denoise_noop" in plan) == (strength == 0.0) and ("denoise_doub" in plan) == (strength > 0.0), plan
    assert "rawnoop" not in plan
    assert "9x7x4x1:f32" in plan, plan           # the gain texture, rgba f32, read by the kernel
    # a list that is not four Bayer gain maps is ignored like the reference does (get_gain_maps_bayer returns 0)
    g2 = api.Graph(cfg_text=_cfg(strength))
    g2.set_source(raw.ctypes.data, api.raw_params(w, h))
    g2.set_dng_opcodes(blobs()["mixed"])
    assert "x4x1:f32" not in "\n".join(l for l in g2.plan().splitlines() if "denoise" in l)
    g.close(); g2.close()


def _oracle_with_gainmap(oracle, raw, planes, strength, w, h, map_os):
    d = oracle.darkroom_defaults(w, h)
    for k in range(3): d.whitebalance[k] = WB[k]
    for k in range(9): d.cam_to_rec2020[k] = CAM[k]
    d.denoise.strength = strength
    gm = np.ascontiguousarray(np.stack(planes, axis=-1).astype(np.float32))
    L = oracle.lib()
    gi = oracle.img(gm)
    L.o_set_gainmap(C.byref(gi), (C.c_float * 4)(*map_os))
    try:
        return oracle.darkroom_run(d, raw)
    finally:
        L.o_set_gainmap(None, None)


@pytest.mark.gpu
@pytest.mark.parametrize("strength", [0.0, 0.4])
def test_gain_maps_end_to_end_from_memory(gpu, oracle, strength):
    w, h = 640, 482
    raw = synth.mosaic(w, h, seed=41)
    planes, blob = gain_maps(h, w)
    pv, ph = planes[0].shape
    # denoise/main.c:188-195: origin and 1 / (spacing * (points - 1)) in fp32
    map_os = [0.0, 0.0, float(np.float32(1.0 / ((1.0 / (ph - 1)) * (ph - 1)))), float(np.float32(1.0 / ((1.0 / (pv - 1)) * (pv - 1))))]
    want = _oracle_with_gainmap(oracle, raw, planes, strength, w, h, map_os)
    plain = _oracle_with_gainmap(oracle, raw, [np.ones_like(p) for p in planes], strength, w, h, map_os)
    assert np.abs(want - plain).max() > 0.01        # the maps do something
    g = gpu.Graph(cfg_text=_cfg(strength))
    buf = np.ascontiguousarray(raw)
    g.set_source(buf.ctypes.data, gpu.raw_params(w, h, wb=WB, cam_to_rec2020=CAM))
    g.set_dng_opcodes(blob)
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    out = np.zeros((oh, ow, 4), dtype=np.float32)
    g.set_sink_buffer(out.ctypes.data, out.nbytes)
    g.run(gpu.RUN_RECORD | gpu.RUN_UPLOAD | gpu.RUN_DOWNLOAD | gpu.RUN_WAIT)
    assert out.shape == want.shape
    assert np.array_equal(out[..., :3], want[..., :3]), float(np.abs(out[..., :3] - want[..., :3]).max())
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("strength", [0.0, 0.4])
def test_gain_maps_end_to_end_from_a_dng_file(gpu, oracle, tmp_path, strength):
    """the file path: OpcodeList2 read from the dng, the stored cfa phase off the canonical one, so that the loader's window
    offset (ox, oy) decides which map belongs to which site (denoise/main.c:26: filter from region.top + oy, region.left + ox)."""
    w, h = 532, 404
    stored = np.array(((1, 2), (0, 1)))                                  # g b / r g: the window starts at (ox, oy) = (0, 1)
    fn = str(tmp_path / "gm.dng")
    x2r = np.array([[1.71665119, -0.35567078, -0.25336628], [-0.66668435, 1.61648124, 0.01576855], [0.01763986, -0.04277061, 0.94210312]])
    cm = tuple((np.linalg.inv(np.array(CAM).reshape(3, 3)) @ x2r).ravel())
    full = np.zeros((h, w), np.uint16)
    synth.write_dng(fn, full, cfa=stored, black=2048, white=15000, neutral=(0.5, 1.0, 2.0 / 3.0), color_matrix=cm)
    p0, ox, oy = gpu.dng_info(fn)
    ww, hh = p0.width, p0.height
    win = synth.mosaic(ww, hh, seed=43)
    full[oy:oy + hh, ox:ox + ww] = win
    # maps in FILE coordinates: the plane whose region starts on (top, left) covers the sites of that parity in the stored image
    planes_file, blob = gain_maps(h, w)
    synth.write_dng(fn, full, cfa=stored, black=2048, white=15000, neutral=(0.5, 1.0, 2.0 / 3.0), color_matrix=cm, opcode_list2=blob)
    # in window coordinates site (y & 1, x & 1) is file site ((y + oy) & 1, (x + ox) & 1): gm[filter] with filter = ((top + oy) & 1) * 2 + ((left + ox) & 1)
    planes = [None] * 4
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        planes[(((dy + oy) & 1) << 1) + ((dx + ox) & 1)] = planes_file[k]
    pv, ph = planes[0].shape
    map_os = [0.0, 0.0, float(np.float32(1.0 / ((1.0 / (ph - 1)) * (ph - 1)))), float(np.float32(1.0 / ((1.0 / (pv - 1)) * (pv - 1))))]
    g = gpu.Graph(cfg_text=_cfg(strength))
    assert g.line("param:i-raw:main:filename:%s" % fn) == 0
    assert g.line("param:i-raw:main:noise a:100.0") == 0 and g.line("param:i-raw:main:noise b:2.0") == 0
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    got = np.zeros((oh, ow, 4), dtype=np.float32)
    g.set_sink_buffer(got.ctypes.data, got.nbytes)
    g.run(gpu.RUN_RECORD | gpu.RUN_UPLOAD | gpu.RUN_DOWNLOAD | gpu.RUN_WAIT)
    d = oracle.darkroom_defaults(ww, hh)
    d.noise_a, d.noise_b = 100.0, 2.0
    d.denoise.strength = strength
    for k in range(4): d.whitebalance[k] = p0.whitebalance[k]
    for k in range(9): d.cam_to_rec2020[k] = p0.cam_to_rec2020[k]
    for k in range(4): d.crop_aabb[k] = p0.crop_aabb[k]
    gm = np.ascontiguousarray(np.stack(planes, axis=-1).astype(np.float32))
    L = oracle.lib()
    gi = oracle.img(gm)
    L.o_set_gainmap(C.byref(gi), (C.c_float * 4)(*map_os))
    try:
        want = oracle.darkroom_run(d, win)
    finally:
        L.o_set_gainmap(None, None)
    assert got.shape == want.shape
    assert np.array_equal(got[..., :3], want[..., :3]), float(np.abs(got[..., :3] - want[..., :3]).max())
    g.close()
