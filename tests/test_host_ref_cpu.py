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


# ---------------------------------------------------------------------------------------------------------------------
# node graphs: what create_nodes of denoise / hilite / demosaic / llap / filmcurv builds, against the REFERENCE's own
# <module>/main.c (tests/golden/host_nodes.json.gz, written by make_golden.py through oracle/ref_nodes_shim.c)
import gzip
import json
import re

NODES = json.loads(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "host_nodes.json.gz")).read())


def _describe(lines, w, h, raw):
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    for ln in lines:
        assert g.line(ln) == 0, ln
    buf = np.zeros((h, w), np.uint16)
    kw = dict(raw)
    for k in ("wb", "crop_aabb"):
        if k in kw:
            kw[k] = tuple(kw[k])
    g.set_source(buf.ctypes.data, api.raw_params(w, h, **kw))
    return g.describe()


def _blocks(text):
    out, cur = {}, None
    for ln in text.splitlines():
        if ln.startswith("module "):
            cur = ln.split()[1]
            out[cur] = []
        if cur is not None and not ln.startswith(" params"):
            out[cur].append(re.sub(r" m=\d+", "", ln))   # request strengths are negotiated by the whole graph, an input of the harness
    return out


def _node_blocks(lines):
    """[(header, [connector lines])] of a module block"""
    nodes = []
    for ln in lines:
        if ln.startswith(" node "):
            nodes.append([ln, []])
        elif ln.startswith("  conn ") and nodes:
            nodes[-1][1].append(ln)
    return nodes


def _documented_deviations(module, ref, case):
    """rewrite the reference's text where the product departs from it ON PURPOSE (each one is stated in DESIGN.md §4)."""
    out, node, dropped = [], "", 0
    for ln in ref:
        if ln.startswith(" node "):
            node = ln.split()[2]
            # 5. colour's sink node writes the as-shot clut temperature back into the `temp` parameter for the gui
            #    (colour/main.c:395-413, :433-434): not built, the nodes behind it move up by one
            if node == "colour:sink":
                dropped += 1
                continue
            if dropped and module == "colour":
                t = ln.split()
                ln = " node %d %s" % (int(t[1]) - dropped, " ".join(t[2:]))
        if node == "colour:sink":
            continue
        # 1. denoise:noop stores one channel: the reference declares rgba and writes (v,0,0,1), every consumer reads .r
        if node == "denoise:noop" and ln.startswith("  conn 1 output:write:rgba:f16"):
            ln = ln.replace("output:write:rgba:f16", "output:write:rggb:f16")
        # 2. rcd_fill: the reference sizes the dispatch in 58x26 px shared memory tiles of 8x8.. threads (demosaic/main.c:125-131);
        #    our kernel tiles internally and the node keeps the image extent
        m = re.match(r" node (\d+) demosaic:rcd_fill (\d+)x(\d+)x1 (.*)", ln)
        if m:
            w, h = case["w"], case["h"]
            if "crop_aabb" in case["raw"]:
                c = case["raw"]["crop_aabb"]
                w, h = c[2] - c[0], c[3] - c[1]
            assert (int(m.group(2)), int(m.group(3))) == ((w + 57) // 58 * 8, (h + 25) // 26 * 8)
            ln = " node %s demosaic:rcd_fill %dx%dx1 %s" % (m.group(1), w, h, m.group(4))
        out.append(ln)
    return out


def test_product_node_graphs_match_reference():
    checked = 0
    for case in NODES:
        mine = _blocks(_describe(case["lines"], case["w"], case["h"], case["raw"]))
        for module, text in case["modules"].items():
            ref = _documented_deviations(module, _blocks(text)[module], case)
            got = mine[module]
            assert ref == got, (case["lines"], case["w"], case["h"], module, [(a, b) for a, b in zip(ref, got) if a != b][:4])
            checked += 1
    assert checked >= 75


def test_xtrans_with_rcd_falls_back_to_gaussian_splats():
    """4. the reference runs RCD, a Bayer algorithm, on any mosaic when method=1 (demosaic/main.c:116); the product keeps method 0 there."""
    b = _blocks(_describe(["param:demosaic:01:method:1"], 516, 390, dict(filters=9)))["demosaic"]
    kernels = [ln.split()[2] for ln in b if ln.startswith(" node ")]
    assert kernels[:4] == ["demosaic:down", "demosaic:gauss", "demosaic:splat", "demosaic:fix"]


def test_live_reference_nodes_random(oracle):
    """the same comparison on random sizes / levels against the compiled reference, where it exists (the build container)."""
    if oracle.ref_host_lib() is None or not hasattr(oracle.ref_host_lib(), "ref_nodes_llap") or not os.path.isdir("/root/reference/src/pipe/modules"):
        pytest.skip("oracle/_ref/libhostref.so or /root/reference not present")
    mg = _make_golden_module()
    rng = np.random.default_rng(20261017)
    for t in range(24):
        xtrans = t % 3 == 2
        blk = 6 if xtrans else 2
        w, h = int(rng.integers(6, 700)) * blk, int(rng.integers(6, 500)) * blk
        lines = []
        if t % 2:
            lines.append("param:denoise:01:strength:%g" % rng.uniform(0.05, 1.0))
        if not xtrans and t % 4 == 0:
            lines.append("param:demosaic:01:method:%d" % rng.integers(1, 3))
        raw = dict(black=float(rng.integers(0, 4096)), white=float(rng.integers(8000, 65535)), wb=(float(rng.uniform(1, 3)), 1.0, float(rng.uniform(1, 3))),
                   noise_a=float(rng.uniform(0.1, 200)), noise_b=float(rng.uniform(0.1, 4)))
        if xtrans:
            raw["filters"] = 9
        case = dict(lines=lines, w=w, h=h, raw=raw)
        mine = _blocks(_describe(lines, w, h, raw))
        for module, text in mg.reference_nodes(lines, w, h, raw).items():
            assert _documented_deviations(module, _blocks(text)[module], case) == mine[module], (case, module)


# ---------------------------------------------------------------------------------------------------------------------
# the whole module pass: config reader, display replacement, roi out / roi in negotiation (sizes AND request strengths),
# create_nodes of every module, repointing through the module layer and commit_params, against the REFERENCE's own graph code
# run over its own bin/default-darkroom.i-raw (tests/golden/host_graph.json.gz, oracle/ref_graph_shim.c)
GRAPHS = json.loads(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "host_graph.json.gz")).read())


def _make_golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


# the reference's bin/default-darkroom.i-mlv ends in llap (no grade): llap feeds the display and the histogram
MLV_CFG = api.DARKROOM_CFG.format(src="i-mlv").replace("connect:llap:01:output:grade:01:input\n", "") \
    .replace("connect:grade:01:output:display:main:input", "connect:llap:01:output:display:main:input") \
    .replace("connect:grade:01:output:hist:01:input", "connect:llap:01:output:hist:01:input")


def _graph_text_product(case):
    if "pfm" in case:
        mg = _make_golden_module()
        fn, _cfg = mg.write_golden_pfm()
        return api.Graph(cfg_text=mg.PFM_CFG % fn).describe().splitlines()
    if "mlv" in case:
        _make_golden_module().write_golden_clip(case["mlv"])     # same pixels, same path as when the golden was made
        g = api.Graph(cfg_text=MLV_CFG)
        for ln in case["lines"]:
            assert g.line(ln) == 0, ln
        return g.describe().splitlines()
    if "luts" in case:
        _make_golden_module().write_golden_luts(clut=case["luts"] == 2)   # same tables, same path as when the golden was made
    mw = mh = 0
    sink, prim, trc = "o-pfm", None, None
    for ln in case["lines"]:    # "#export:max:<w>:<h>": the cli's --width / --height (a resize module in front of the sink)
        if ln.startswith("#export:max:"):
            mw, mh = [int(x) for x in ln.split(":")[2:4]]
        if ln.startswith("#export:colour:"):   # --colour-prim / --colour-trc
            prim, trc = [int(x) for x in ln.split(":")[2:4]]
        if ln.startswith("#export:sink:"):     # --format
            sink = ln.split(":")[2]
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"), sink=sink, prim=prim, trc=trc, max_width=mw, max_height=mh)
    assert g.line("param:i-raw:main:filename:none.raw") == 0   # the last line of the reference's default-darkroom.i-raw
    for ln in case["lines"]:
        if not ln.startswith("#"):
            assert g.line(ln) == 0, ln
    buf = np.zeros((case["h"], case["w"]), np.uint16)
    kw = dict(case["raw"])
    for k in ("wb", "crop_aabb"):
        if k in kw:
            kw[k] = tuple(kw[k])
    g.set_source(buf.ctypes.data, api.raw_params(case["w"], case["h"], **kw))
    return g.describe().splitlines()


def _graph_text_reference(case, text):
    """the reference's text with the documented deviations applied (see _documented_deviations)"""
    out = text.splitlines()
    fixed, start = [], 0
    for i in range(len(out) + 1):   # per module, for the rewrites that need to know the node
        if i == len(out) or (out[i].startswith("module ") and i > start):
            fixed += _documented_deviations(out[start].split()[1], out[start:i], case)
            start = i
    return fixed


def test_product_module_pass_matches_reference_graph_code():
    for case in GRAPHS:
        ref = _graph_text_reference(case, case["text"])
        got = _graph_text_product(case)
        assert len(ref) == len(got), (case["lines"], case["w"], case["h"], len(ref), len(got))
        bad = [(a[:200], b[:200]) for a, b in zip(ref, got) if a != b]
        assert not bad, (case["lines"], case["w"], case["h"], bad[:3])
    assert len(GRAPHS) >= 37 and sum("mlv" in c for c in GRAPHS) == 2 and sum("pfm" in c for c in GRAPHS) == 1 and sum("luts" in c for c in GRAPHS) == 2
    assert sum(any(l.startswith("#export:max") for l in c["lines"]) for c in GRAPHS) == 5 and sum(any(l.startswith("feedback:") for l in c["lines"]) for c in GRAPHS) == 2
    assert sum(any(l.startswith("#export:colour") for l in c["lines"]) for c in GRAPHS) == 4 and sum(any(l.startswith("#export:sink:o-jpg") for l in c["lines"]) for c in GRAPHS) == 3


def test_live_reference_graph_random(oracle):
    """the same on random sizes and parameters against the compiled reference, where it exists (the build container)."""
    if oracle.ref_host_lib() is None or not hasattr(oracle.ref_host_lib(), "ref_graph_describe") or not os.path.isdir("/root/reference/src/pipe/modules"):
        pytest.skip("oracle/_ref/libhostref.so or /root/reference not present")
    rng = np.random.default_rng(1017)
    for t in range(30):
        xtrans = t % 3 == 2
        blk = 6 if xtrans else 2
        w, h = int(rng.integers(6, 900)) * blk, int(rng.integers(6, 600)) * blk
        lines = []
        if t % 2:
            lines.append("param:denoise:01:strength:%g" % rng.uniform(0.05, 1.0))
        if not xtrans and t % 4 == 0:
            lines.append("param:demosaic:01:method:%d" % rng.integers(1, 3))
        if t % 5 == 1:
            lines.append("param:crop:01:rotate:%g" % rng.choice([0.0, 90.0, 180.0, 270.0, rng.uniform(-20, 20)]))
        if t % 5 == 3:
            a, b = sorted(rng.uniform(0, 1, 2)); c, d = sorted(rng.uniform(0, 1, 2))
            lines.append("param:crop:01:crop:%g:%g:%g:%g" % (a, b + 0.05, c, d + 0.05))
        if t % 7 == 2:
            lines += ["param:colour:01:exposure:%g" % rng.uniform(-2, 2), "param:colour:01:matrix:%d" % rng.integers(0, 3), "param:colour:01:temp:%g" % rng.uniform(2500, 9000)]
        raw = dict(black=float(rng.integers(0, 4096)), white=float(rng.integers(8000, 65535)), wb=(float(rng.uniform(1, 3)), 1.0, float(rng.uniform(1, 3))),
                   noise_a=float(rng.uniform(0.1, 200)), noise_b=float(rng.uniform(0.1, 4)))
        if xtrans:
            raw["filters"] = 9
        case = dict(lines=lines, w=w, h=h, raw=raw)
        ref = _graph_text_reference(case, oracle.ref_graph_describe(w, h, lines, raw))
        got = _graph_text_product(case)
        bad = [(a[:200], b[:200]) for a, b in zip(ref, got) if a != b]
        assert len(ref) == len(got) and not bad, (case, len(ref), len(got), bad[:3])


def test_live_reference_graph_with_random_luts(oracle, tmp_path):
    """lut tables of random shapes (bands, channel counts, f16 / f32) wired to colour in random combinations, random temperatures
    incl. as shot: i-lut's roi and formats, colour's connectors, nodes, push constants and committed block against the compiled
    reference (its i-lut/main.c, colour/main.c)."""
    if oracle.ref_host_lib() is None or not hasattr(oracle.ref_host_lib(), "ref_graph_describe") or not os.path.isdir("/root/reference/src/pipe/modules"):
        pytest.skip("oracle/_ref/libhostref.so or /root/reference not present")
    import struct
    rng = np.random.default_rng(2024)
    for t in range(16):
        w, h = int(rng.integers(40, 700)) * 2, int(rng.integers(40, 500)) * 2
        lines = []
        use = [n for n, pr in (("clut", 0.6), ("abney", 0.7), ("spectra", 0.7)) if rng.uniform() < pr]
        for name in use:
            if name == "clut":
                ht = int(rng.integers(8, 48)); wd = ht * int(rng.choice([3, 6, 9, 12])); ch, dt = 2, np.float16
            elif name == "abney":
                wd, ht, ch, dt = int(rng.integers(16, 128)), int(rng.integers(16, 96)), 2, np.float16
            else:
                wd, ht, ch, dt = int(rng.integers(16, 96)), int(rng.integers(16, 96)), 4, rng.choice([np.float16, np.float32])
            fn = str(tmp_path / ("%s_%d.lut" % (name, t)))
            with open(fn, "wb") as f:
                f.write(struct.pack("<IHBBII", 1234, 2, ch, 0 if dt == np.float16 else 1, wd, ht))
                f.write(rng.uniform(0.1, 0.9, (ht, wd, ch)).astype(dt).tobytes())
            lines += ["module:i-lut:%s" % name, "param:i-lut:%s:filename:%s" % (name, fn), "connect:i-lut:%s:output:colour:01:%s" % (name, name)]
        lines += ["param:colour:01:matrix:%d" % rng.choice([1, 4]), "param:colour:01:temp:%g" % rng.choice([0.0, rng.uniform(1500, 16000)]),
                  "param:colour:01:gamut:%d" % rng.integers(0, 4), "param:colour:01:sat:%g" % rng.uniform(0.5, 1.5)]
        raw = dict(wb=(float(rng.uniform(1, 3)), 1.0, float(rng.uniform(1, 3))))
        case = dict(lines=lines, w=w, h=h, raw=raw)
        ref = _graph_text_reference(case, oracle.ref_graph_describe(w, h, lines, raw))
        got = _graph_text_product(case)
        bad = [(a[:200], b[:200]) for a, b in zip(ref, got) if a != b]
        assert len(ref) == len(got) and not bad, (case, len(ref), len(got), bad[:3])


def test_live_reference_graph_with_feedback_edges(oracle):
    """`feedback:` connections through the whole module pass (traversal incl. the second round of graph-traverse.inc:131-148,
    roi negotiation, create_nodes, repointing) against the compiled reference."""
    if oracle.ref_host_lib() is None or not hasattr(oracle.ref_host_lib(), "ref_graph_describe") or not os.path.isdir("/root/reference/src/pipe/modules"):
        pytest.skip("oracle/_ref/libhostref.so or /root/reference not present")
    raw = dict(black=1024.0, white=15000.0, wb=(2.0, 1.0, 1.5), noise_a=10.0, noise_b=1.0)
    for lines, w, h in ((["feedback:grade:01:output:colour:01:spectra"], 640, 480),
                        (["feedback:llap:01:output:colour:01:spectra", "param:denoise:01:strength:0.3"], 1002, 668),
                        (["feedback:crop:01:output:colour:01:spectra", "param:crop:01:rotate:90"], 322, 246)):
        case = dict(lines=lines, w=w, h=h, raw=raw)
        ref = _graph_text_reference(case, oracle.ref_graph_describe(w, h, lines, raw))
        got = _graph_text_product(case)
        bad = [(a[:200], b[:200]) for a, b in zip(ref, got) if a != b]
        assert len(ref) == len(got) and not bad, (case, len(ref), len(got), bad[:3])


# ---------------------------------------------------------------------------------------------------------------------
# the config grammar against the REFERENCE's own graph-io.c: return code per line (0 ok, > 0 warning, < 0 fatal) and the
# state the lines leave behind (tests/golden/host_cfg.json.gz)
CFG = json.loads(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "host_cfg.json.gz")).read())


def _product_cfg_lines(lines):
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw") + "param:i-raw:main:filename:none.raw\n", sink=None)
    return [g.line(ln) for ln in lines], g.state()


def test_config_lines_match_reference_reader():
    codes, state = _product_cfg_lines(CFG["lines"])
    assert codes == CFG["codes"], [(ln, a, b) for ln, a, b in zip(CFG["lines"], CFG["codes"], codes) if a != b]
    assert state.splitlines() == CFG["state"].splitlines()
    assert 12 in codes and 4 in codes and 2 in codes and 1 in codes     # cycle, index out of range, no such parameter, no such module


def test_keyframe_and_feedback_lines_are_read():
    """like the reference's reader (graph-io.c:243-249) both are accepted; tests/test_keyframes_cpu.py covers what they do."""
    codes, _ = _product_cfg_lines(["module:grade:02", "keyframe:3:colour:01:exposure:0:1:1.5", "feedback:grade:01:output:grade:02:input",
                                   "keyframe:3:nosuch:01:exposure:0:1:1.5", "keyframe:3:colour:01:nosuch:0:1:1.5", "keyframe:3:colour:01:exposure:0:5:1.5"])
    assert codes == [0, 0, 0, 1, 2, 4]


def test_live_config_line_fuzz(oracle):
    if oracle.ref_host_lib() is None or not hasattr(oracle.ref_host_lib(), "ref_config_lines") or not os.path.isdir("/root/reference/src/pipe/modules"):
        pytest.skip("oracle/_ref/libhostref.so or /root/reference not present")
    rng = np.random.default_rng(77)
    mods = ["colour:01", "crop:01", "llap:01", "grade:01", "denoise:01", "hilite:01", "demosaic:01", "filmcurv:01", "nosuch:01", "colour:02"]
    pars = ["exposure", "crop", "sigma", "gain", "strength", "white", "method", "light", "mat", "rbmap", "rotate", "perspect", "nosuch", "clarity", "cnt"]
    vals = ["0.5", "-1", "3", "1e-3", "abc", "", "7.25", "0", "2", "1e9"]
    lines = []
    for t in range(400):
        kind = rng.integers(0, 6)
        if kind < 3:
            n = int(rng.integers(0, 6))
            lines.append("param:%s:%s" % (rng.choice(mods), rng.choice(pars)) + "".join(":" + rng.choice(vals) for _ in range(n)))
        elif kind == 3:
            lines.append("paramsub:%s:%s:%d" % (rng.choice(mods), rng.choice(pars), rng.integers(0, 12)) + "".join(":" + rng.choice(vals) for _ in range(int(rng.integers(1, 4)))))
        elif kind == 4:
            a, b = rng.choice(mods[:8], 2)
            lines.append("connect:%s:output:%s:input" % (a, b))
        else:
            lines.append("module:%s:%02d" % (rng.choice(["grade", "llap", "colour", "crop", "nosuch"]), rng.integers(1, 4)))
    ref_codes, ref_state = oracle.ref_config_lines(lines)
    codes, state = _product_cfg_lines(lines)
    assert codes == ref_codes, [(ln, a, b) for ln, a, b in zip(lines, ref_codes, codes) if a != b][:5]
    assert state.splitlines() == ref_state.splitlines(), [(a[:160], b[:160]) for a, b in zip(ref_state.splitlines(), state.splitlines()) if a != b][:3]


def test_pfm_writer_matches_reference_live(oracle, tmp_path):
    """o-pfm: the oracle's writer against the reference's own write_sink (o-pfm/main.c compiled in place), byte for byte, over
    sizes whose digit counts move the header padding (payload 16 byte aligned); the product's sink is checked against the
    oracle's writer on the GPU."""
    href = oracle.ref_host_lib()
    if href is None or not hasattr(href, "ref_write_pfm"):
        pytest.skip("oracle/_ref/libhostref.so not present")
    rng = np.random.default_rng(3)
    for w, h in ((5, 3), (9, 9), (10, 9), (99, 100), (100, 100), (1000, 7), (1234, 3), (12345, 1)):
        rgba = rng.random((h, w, 4), dtype=np.float32)
        a, b = str(tmp_path / ("o_%d_%d" % (w, h))), str(tmp_path / ("r_%d_%d" % (w, h)))
        assert oracle.lib().o_write_pfm((a + ".pfm").encode(), oracle.fptr(rgba), w, h) == 0
        href.ref_write_pfm(b.encode(), oracle.fptr(rgba), w, h)
        da, db = open(a + ".pfm", "rb").read(), open(b + ".pfm", "rb").read()
        assert da == db, (w, h, da[:40], db[:40])


def test_imlv_known_camera_matrix_from_the_checkout(oracle):
    """a clip of a camera that IS in dcraw's adobe_coeff table (i-mlv/main.c:165-201): with the vkdt checkout named as basedir the
    product reads that table from the checkout's own file and has to hand on the reference's image parameters (white balance from
    the matrix, cam_to_rec2020) bit for bit: compared through the whole module pass with the reference's own i-mlv/main.c."""
    if oracle.ref_host_lib() is None or not hasattr(oracle.ref_host_lib(), "ref_graph_describe") or not os.path.isdir("/root/reference/src/pipe/modules"):
        pytest.skip("oracle/_ref/libhostref.so or /root/reference not present")
    mg = _make_golden_module()
    os.makedirs(mg.MLV_DIR, exist_ok=True)
    for camera in ("Canon EOS 5D Mark III", "canon eos 7d", "Nikon D850"):          # the lookup ignores case
        fn = os.path.join(mg.MLV_DIR, "known.mlv")
        pix = mg.synth.mosaic(256, 192, seed=5)
        mg.synth.write_mlv(fn, [pix], bpp=14, black=2048, white=15000, camera_name=camera)
        lines = ["param:i-mlv:main:filename:" + fn]
        ref = _graph_text_reference(dict(lines=lines, mlv="known"), oracle.ref_graph_describe(256, 192, lines, {}, cfg="bin/default-darkroom.i-mlv"))
        try:
            api.set_basedir("/root/reference/src/pipe")
            g = api.Graph(cfg_text=MLV_CFG)
            for ln in lines:
                assert g.line(ln) == 0, ln
            got = g.describe().splitlines()
            g.close()
        finally:
            api.set_basedir("")
        bad = [(a[:240], b[:240]) for a, b in zip(ref, got) if a != b]
        assert len(ref) == len(got) and not bad, (camera, bad[:3])
        # and it is not the identity branch
        img = [ln for ln in got if ln.startswith("module i-mlv")][0]
        assert "1.71665" not in img or camera == "", img[:200]
