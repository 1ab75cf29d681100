"""crop and colour: the parameter blocks the kernels are handed, pinned against the REFERENCE's own host code.
tests/golden/host_ref.npz holds random parameter sets with the output of crop/main.c (modify_roi_out + commit_params) and
colour/main.c (commit_params incl. the CAT16 white point and the RBF solve), compiled in place from /root/reference by
tests/golden/make_golden.py.  Both the oracle's restatement and the product's module callbacks have to reproduce them
bit for bit (NaNs, which only unconnected `picked` inputs produce, compare equal)."""
import ctypes as C
import os
import numpy as np
import pytest

from vkdt_b200 import api

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "host_ref.npz"))


def _same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0))


def test_oracle_crop_matches_reference(oracle):
    for inp, out in zip(G["crop_in"], G["crop_out"]):
        ori, w, h = int(inp[0]), int(inp[1]), int(inp[2])
        persp = [float(np.float32(x)) for x in inp[3:11]]
        ow, oh, f = oracle.crop_oracle(ori, w, h, persp, [float(x) for x in inp[11:15]], float(inp[15]))
        assert (ow, oh) == (int(out[0]), int(out[1])) and np.array_equal(f.view(np.uint32), out[2:].astype(np.float32).view(np.uint32)), inp


def test_oracle_colour_commit_matches_reference(oracle):
    for par, img, out in zip(G["colour_params"], G["colour_img"], G["colour_out"]):
        f, wbo = oracle.colour_commit_oracle(par.tobytes(), list(img[:4]), list(img[4:13]), int(img[13]), int(img[14]))
        assert _same(f, out[:242]) and _same(wbo, out[242:246])


def _graph(w, h, **kw):
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    raw = np.zeros((h, w), np.uint16)
    g._keep.append(raw)
    return g, raw, kw


def test_product_crop_matches_reference():
    """the product's crop module (roi + committed homography / rotation / window) through the C-ABI, no GPU."""
    checked = 0
    for inp, out in zip(G["crop_in"][:150], G["crop_out"][:150]):
        ori, w, h = int(inp[0]), int(inp[1]), int(inp[2])
        w2, h2 = (w // 2) * 2, (h // 2) * 2           # i-raw rounds the mosaic to the cfa block before crop sees it
        if (w2, h2) != (w, h) or w < 64 or h < 64:
            continue
        g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
        assert g.line("param:crop:01:perspect:" + ":".join(repr(float(np.float32(x))) for x in inp[3:11])) == 0
        assert g.line("param:crop:01:crop:" + ":".join(repr(float(x)) for x in inp[11:15])) == 0
        assert g.line("param:crop:01:rotate:" + repr(float(inp[15]))) == 0
        raw = np.zeros((h, w), np.uint16)
        rp = api.raw_params(w, h)
        rp.orientation = ori
        g.set_source(raw.ctypes.data, rp)
        try:
            f = g.committed_params("crop")
        except api.VkbError:
            assert int(out[0]) == 0 or int(out[1]) == 0    # a window without area: refused
            continue
        ow, oh = [int(x) for x in [l for l in g.plan().splitlines() if l.startswith("sink")][0].split()[2].split("x")]
        assert (ow, oh) == (int(out[0]), int(out[1])), inp
        assert np.array_equal(f.view(np.uint32), out[2:].astype(np.float32).view(np.uint32)), inp
        checked += 1
    assert checked >= 30


def test_product_colour_commit_matches_reference(oracle):
    """the product's colour module: 242 committed floats for i-raw style input (custom primaries, linear)."""
    names = [("exposure", "f", 1), ("sat", "f", 1), ("picked", "i", 1), ("matrix", "i", 1), ("gamut", "i", 1), ("clip", "i", 1), ("clipmax", "f", 1),
             ("temp", "f", 1), ("white", "f", 4), ("mat", "f", 9), ("mode", "i", 1), ("cnt", "i", 1), ("rbmap", "f", 144)]
    checked = 0
    for par, img, out in zip(G["colour_params"], G["colour_img"], G["colour_out"]):
        if int(img[13]) != 0 or int(img[14]) != 0:
            continue                                   # other primaries / trcs reach colour only behind other sources
        g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
        raw_par = par.tobytes()
        off = 0
        for name, kind, cnt in names:
            vals = np.frombuffer(raw_par, np.float32 if kind == "f" else np.int32, cnt, off)
            off += 4 * cnt
            assert g.line("param:colour:01:%s:" % name + ":".join(repr(float(v)) if kind == "f" else str(int(v)) for v in vals)) == 0
        raw = np.zeros((64, 64), np.uint16)
        g.set_source(raw.ctypes.data, api.raw_params(64, 64, wb=tuple(float(x) for x in img[:3]), cam_to_rec2020=tuple(float(x) for x in img[4:13])))
        f = g.committed_params("colour")
        assert _same(f, out[:242]), (checked, np.nonzero(np.nan_to_num(f, nan=-7.0) != np.nan_to_num(out[:242], nan=-7.0))[0][:8])
        checked += 1
    assert checked >= 60
