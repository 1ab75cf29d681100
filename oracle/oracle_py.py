"""ORACLE — test infrastructure only.

ctypes bindings for oracle/liboracle.so (CPU restatement of vkdt's raw->display kernels) and
oracle/_ref/libmlvref.so (the reference's own MLV decoder compiled in place).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import struct
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OImg(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("c", C.c_int), ("p", C.POINTER(C.c_float))]


class DenoiseParams(C.Structure):
    _fields_ = [("strength", C.c_float), ("luma", C.c_float), ("detail", C.c_float), ("pad", C.c_float),
                ("edges", C.c_float * 4), ("gainmap", C.c_int)]


class HiliteParams(C.Structure):
    _fields_ = [("white", C.c_float), ("desat", C.c_float), ("soft", C.c_float)]


class DemosaicParams(C.Structure):
    _fields_ = [("colour", C.c_int), ("method", C.c_int)]


class CropParams(C.Structure):
    _fields_ = [("perspect", C.c_float * 8), ("crop", C.c_float * 4), ("rotate", C.c_float)]


class ColourParams(C.Structure):
    _fields_ = [("exposure", C.c_float), ("sat", C.c_float), ("picked", C.c_int), ("matrix", C.c_int),
                ("gamut", C.c_int), ("clip", C.c_int), ("clipmax", C.c_float), ("temp", C.c_float),
                ("white", C.c_float * 4), ("mat", C.c_float * 9), ("mode", C.c_int), ("cnt", C.c_int),
                ("rbmap", C.c_float * 144), ("import_", C.c_char * 8)]


class FilmcurvParams(C.Structure):
    _fields_ = [("light", C.c_float), ("contrast", C.c_float), ("bias", C.c_float), ("colour", C.c_int),
                ("chroma", C.c_float), ("rolloff", C.c_float), ("red", C.c_float), ("yellow", C.c_float),
                ("blue", C.c_float), ("shadows", C.c_float)]


class LlapParams(C.Structure):
    _fields_ = [("sigma", C.c_float), ("shadows", C.c_float), ("hilights", C.c_float), ("clarity", C.c_float)]


class GradeParams(C.Structure):
    _fields_ = [("lift", C.c_float * 4), ("gamma", C.c_float * 4), ("gain", C.c_float * 4),
                ("offset", C.c_float * 4), ("mode", C.c_int), ("sh_pivot", C.c_float), ("hi_pivot", C.c_float)]


class Darkroom(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("filters", C.c_uint32),
                ("crop_aabb", C.c_uint32 * 4), ("black", C.c_float * 4), ("white", C.c_float * 4),
                ("whitebalance", C.c_float * 4), ("cam_to_rec2020", C.c_float * 9),
                ("noise_a", C.c_float), ("noise_b", C.c_float), ("orientation", C.c_uint32),
                ("colour_primaries", C.c_int), ("colour_trc", C.c_int),
                ("denoise", DenoiseParams), ("hilite", HiliteParams), ("demosaic", DemosaicParams),
                ("crop", CropParams), ("colour", ColourParams), ("filmcurv", FilmcurvParams),
                ("llap", LlapParams), ("grade", GradeParams),
                ("enable_llap", C.c_int), ("enable_grade", C.c_int),
                ("enable_colenc", C.c_int), ("colenc_prim", C.c_int), ("colenc_trc", C.c_int), ("sink_unorm8", C.c_int)]


_lib = None
_ref = None


def build(ref=True):
    """compile liboracle.so (and oracle/_ref when /root/reference is present). building the checker is not using it."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    if ref and os.path.isdir("/root/reference/src/pipe/modules/i-mlv"):
        subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True)


def set_threads(n=None):
    """OpenMP threads of the restatement; returns the count in effect.  torchrun exports OMP_NUM_THREADS=1 to its
    workers, which would silently turn the "all host cores" CPU baseline into a single-threaded one."""
    lib()
    gomp = C.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(int(n or os.cpu_count() or 1))
    return int(gomp.omp_get_max_threads())


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _lib = C.CDLL(path)
        _lib.o_darkroom_run.restype = C.c_int
    return _lib


def ref_lib():
    """the reference's own video_mlv.c, compiled by `make -C oracle ref`; None when it was never built."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libmlvref.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
    return _ref


_href = None


def ref_host_lib():
    """the HOST side of the reference's crop/main.c and colour/main.c, compiled in place by `make -C oracle ref`
    (oracle/ref_crop_shim.c, ref_colour_shim.c); None when it was never built."""
    global _href
    if _href is None:
        path = os.path.join(_HERE, "_ref", "libhostref.so")
        if not os.path.exists(path):
            return None
        _href = C.CDLL(path)
    return _href


def ref_crop(orientation, in_w, in_h, perspect, crop, rotate):
    """the reference's crop modify_roi_out + commit_params: (out_w, out_h, committed[20])."""
    ow, oh, f = C.c_uint32(), C.c_uint32(), (C.c_float * 20)()
    ref_host_lib().ref_crop(C.c_uint32(orientation), C.c_uint32(in_w), C.c_uint32(in_h), (C.c_float * 8)(*perspect), (C.c_float * 4)(*crop),
                            C.c_float(rotate), C.byref(ow), C.byref(oh), f)
    return ow.value, oh.value, np.array(list(f), dtype=np.float32)


def crop_oracle(orientation, in_w, in_h, perspect, crop, rotate):
    """the oracle's restatement of the same: (out_w, out_h, committed[20])."""
    ow, oh, f = C.c_uint32(), C.c_uint32(), np.zeros(20, np.float32)
    cr, rot = (C.c_float * 4)(*crop), C.c_float(rotate)
    lib().o_crop_roi_out(C.c_uint32(orientation), C.c_uint32(in_w), C.c_uint32(in_h), cr, C.byref(rot), C.byref(ow), C.byref(oh))
    lib().o_crop_commit(C.c_uint32(orientation), C.c_uint32(in_w), C.c_uint32(in_h), (C.c_float * 8)(*perspect), cr, C.byref(rot), fptr(f))
    return ow.value, oh.value, f


def ref_colour_commit(params_bytes, img_wb, cam_to_rec2020, primaries, trc):
    """the reference's colour commit_params on a raw parameter block: (committed[242], white written back[4])."""
    f, wbo = (C.c_float * 256)(), (C.c_float * 4)()
    n = ref_host_lib().ref_colour_commit(params_bytes, C.c_uint32(len(params_bytes)), (C.c_float * 4)(*img_wb), (C.c_float * 9)(*cam_to_rec2020),
                                         C.c_int(primaries), C.c_int(trc), wbo, f)
    return np.array(list(f), dtype=np.float32)[:n], np.array(list(wbo), dtype=np.float32)


def colour_commit_oracle(params_bytes, img_wb, cam_to_rec2020, primaries, trc):
    p = ColourParams.from_buffer_copy(params_bytes)
    pwb = (C.c_float * 4)(*[p.white[k] for k in range(4)])
    f = np.zeros(256, np.float32)
    lib().o_colour_commit(C.byref(p), pwb, (C.c_float * 4)(*img_wb), (C.c_float * 9)(*cam_to_rec2020), C.c_int(primaries), C.c_int(trc), fptr(f))
    return f[:242], np.array(list(pwb), dtype=np.float32)


def img(a):
    """wrap a float32 numpy array (h,w) or (h,w,4) as oimg_t (keeps a reference to the array)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    h, w = a.shape[:2]
    c = 1 if a.ndim == 2 else a.shape[2]
    o = OImg(w, h, c, a.ctypes.data_as(C.POINTER(C.c_float)))
    o._keep = a
    return o


def new_img(h, w, c=1):
    a = np.zeros((h, w) if c == 1 else (h, w, c), dtype=np.float32)
    return a, img(a)


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def f4(*v):
    return (C.c_float * 4)(*v)


def i4(*v):
    return (C.c_int * 4)(*v)


def darkroom_defaults(width, height):
    d = Darkroom()
    lib().o_darkroom_defaults(C.byref(d), C.c_uint32(width), C.c_uint32(height))
    return d


def darkroom_out_size(d):
    w, h = C.c_uint32(), C.c_uint32()
    lib().o_darkroom_out_size(C.byref(d), C.byref(w), C.byref(h))
    return w.value, h.value


_STAGE_CH = {1: 1, 2: 1, 3: 4, 4: 4, 5: 4, 6: 4, 7: 4}


def darkroom_run(d, raw, stage=-1):
    """run the default darkroom graph on a (H,W) uint16 mosaic; returns (oh,ow,4) float32, or the
    intermediate image of `stage` (see vkdt_oracle.h)."""
    raw = np.ascontiguousarray(raw, dtype=np.uint16)
    assert raw.shape == (d.height, d.width)
    ow, oh = darkroom_out_size(d)
    cw, ch = d.crop_aabb[2] - d.crop_aabb[0], d.crop_aabb[3] - d.crop_aabb[1]
    rp = raw.ctypes.data_as(C.POINTER(C.c_uint16))
    if stage < 0:
        out = np.zeros((oh, ow, 4), dtype=np.float32)
        lib().o_darkroom_run(C.byref(d), rp, fptr(out), -1, None)
        return out
    c = _STAGE_CH[stage]
    shape = (ch, cw) if stage <= 2 else ((ch, cw, 4) if stage == 3 else (oh, ow, 4))
    so = np.zeros(shape, dtype=np.float32)
    lib().o_darkroom_run(C.byref(d), rp, None, stage, fptr(so))
    return so


def mlv_unpack(words, pixel_cnt, bpp):
    """words: uint16 array of the packed payload with at least one spare word at the end."""
    words = np.ascontiguousarray(words, dtype=np.uint16)
    out = np.zeros(pixel_cnt, dtype=np.uint16)
    lib().o_mlv_unpack(words.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_uint64(pixel_cnt), C.c_int(bpp),
                       out.ctypes.data_as(C.POINTER(C.c_uint16)))
    return out


def ref_mlv_decode(filename, frame=0):
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref/libmlvref.so not built (run `make -C oracle ref` where /root/reference exists)")
    info = (C.c_int * 6)()
    if r.ref_mlv_info(filename.encode(), info):
        raise RuntimeError("reference mlv_open_clip failed on " + filename)
    w, h = info[0], info[1]
    out = np.zeros((h, w), dtype=np.uint16)
    if r.ref_mlv_decode(filename.encode(), C.c_uint64(frame), out.ctypes.data_as(C.POINTER(C.c_uint16))):
        raise RuntimeError("reference mlv_get_frame failed")
    return out, dict(width=w, height=h, bpp=info[2], black=info[3], white=info[4], frames=info[5])


class RefNodesIn(C.Structure):
    _fields_ = [("in_full_wd", C.c_uint32), ("in_full_ht", C.c_uint32), ("filters", C.c_uint32), ("black", C.c_float * 4), ("white", C.c_float * 4),
                ("wb", C.c_float * 4), ("crop_aabb", C.c_uint32 * 4), ("noise_a", C.c_float), ("noise_b", C.c_float), ("in_chan", C.c_char_p),
                ("in_format", C.c_char_p), ("out_chan", C.c_char_p), ("out_format", C.c_char_p), ("out_marker", C.c_uint32), ("moddir", C.c_char_p), ("param", C.c_char_p), ("param_size", C.c_uint32)]


def parse_described_modules(text):
    """split the text of vkb_graph_describe / ref_nodes_* into {module name: block} (blocks keep their lines)."""
    blocks, cur = {}, None
    for ln in text.splitlines():
        if ln.startswith("module "):
            cur = ln.split()[1]
            blocks[cur] = []
        if cur is not None:
            blocks[cur].append(ln)
    return blocks


def _img_fields(module_line):
    """image parameters of a `module` line as the bits the reference harness takes."""
    kv = dict(t.split("=") for t in module_line.split() if "=" in t)
    f32 = lambda s: [float(np.array([int(x, 16)], np.uint32).view(np.float32)[0]) for x in s.split(",")]
    return dict(filters=int(kv["filters"]), black=f32(kv["black"]), white=f32(kv["white"]), wb=f32(kv["wb"]), crop=[int(x) for x in kv["crop"].split(",")], noise=f32(kv["noise"]))


def ref_nodes(module, prev_block, block, refdir="/root/reference", img_line=None):
    """run the reference's own <module>/main.c (roi callbacks + create_nodes) on what the product's describe() says enters the
    module: `prev_block` = lines of the module feeding it (image parameters, output connector), `block` = the module's own lines
    (parameter block, negotiated output channels / format / request strength).  `img_line`: the image parameters the module
    copies in its roi pass when they are not the feeding module's final ones (a reference `imgout` line).
    returns (the reference's text for the module, its imgout line)."""
    ip = _img_fields(img_line or prev_block[0])
    prev_out = [ln.split() for ln in prev_block if ln.startswith(" mconn") and ln.split()[2].split(":")[1] in ("write", "source")][0]
    _, _, pchan, pfmt = prev_out[2].split(":")
    pfw, pfh = prev_out[3][4:].split("/")[0].split("x")
    out = [ln.split() for ln in block if ln.startswith(" mconn") and ln.split()[2].startswith("output:")][0]
    blob = bytes.fromhex([ln for ln in block if ln.startswith(" params")][0].split()[1] if len([ln for ln in block if ln.startswith(" params")][0].split()) > 1 else "")
    a = RefNodesIn(int(pfw), int(pfh), ip["filters"], (C.c_float * 4)(*ip["black"]), (C.c_float * 4)(*ip["white"]), (C.c_float * 4)(*ip["wb"]),
                   (C.c_uint32 * 4)(*ip["crop"]), ip["noise"][0], ip["noise"][1], pchan.encode(), pfmt.encode(), out[2].split(":")[2].encode(), out[2].split(":")[3].encode(), int(out[4][2:]),
                   os.path.join(refdir, "src/pipe/modules", module).encode(), blob, len(blob))
    buf = C.create_string_buffer(1 << 18)
    fn = getattr(ref_host_lib(), "ref_nodes_" + module)
    fn.restype = C.c_int
    n = fn(C.byref(a), buf, len(buf))
    assert n > 0, "ref_nodes_%s failed: %d" % (module, n)
    text = buf.value.decode()
    first, rest = text.split("\n", 1)
    assert first.startswith("imgout ")
    return rest, first


def ref_graph_describe(w, h, lines=(), raw=None, sink="o-pfm", refdir="/root/reference", cfg="bin/default-darkroom.i-raw"):
    """the reference's own module pass (config reader, dt_graph_replace_display, roi out / roi in / create nodes, commit_params;
    oracle/ref_graph_shim.c) over one of its .cfg files with a w x h source: text in the form of vkb_graph_describe."""
    raw = dict(raw or {})
    wb = list(raw.get("wb", (1.0, 1.0, 1.0))) + [1.0]
    crop = raw.get("crop_aabb") or (0, 0, w, h)
    a = RefNodesIn(w, h, raw.get("filters", 0x5d5d5d5d), (C.c_float * 4)(*[raw.get("black", 2048.0)] * 4), (C.c_float * 4)(*[raw.get("white", 15000.0)] * 4),
                   (C.c_float * 4)(*wb[:4]), (C.c_uint32 * 4)(*crop), raw.get("noise_a", 1.0), raw.get("noise_b", 1.0), b"", b"", b"", b"", 0, b"", b"", 0)
    buf = C.create_string_buffer(1 << 21)
    fn = ref_host_lib().ref_graph_describe
    fn.restype = C.c_int
    n = fn(os.path.join(refdir, "src/pipe").encode(), os.path.join(refdir, cfg).encode(), "\n".join(lines).encode(), sink.encode(), C.byref(a), buf, len(buf))
    assert n > 0, "ref_graph_describe failed: %d" % n
    return buf.value.decode()


def ref_config_lines(lines, refdir="/root/reference", cfg="bin/default-darkroom.i-raw"):
    """return codes of the reference's dt_graph_read_config_line for each line on top of `cfg`, and the resulting state text."""
    codes = (C.c_int * max(1, len(lines)))()
    buf = C.create_string_buffer(1 << 20)
    fn = ref_host_lib().ref_config_lines
    fn.restype = C.c_int
    n = fn(os.path.join(refdir, "src/pipe").encode(), os.path.join(refdir, cfg).encode(), "\n".join(lines).encode(), codes, len(lines), buf, len(buf))
    assert n == len(lines), "ref_config_lines failed: %d" % n
    return list(codes)[:n], buf.value.decode()


# ---- the reference's own compute shaders, compiled as C++ (oracle/glsl/comp2cpp.py + glsl_shim.h -> oracle/_ref/libshaderref.so) ----
class ShaderImage(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("wd", C.c_int), ("ht", C.c_int), ("chan", C.c_int), ("f16", C.c_int)]


_shader = None


def ref_shader_lib():
    global _shader
    if _shader is None:
        path = os.path.join(_HERE, "_ref", "libshaderref.so")
        if not os.path.exists(path):
            return None
        _shader = C.CDLL(path)
    return _shader


def ref_shader(module, kernel, params, push, bindings, wd, ht, dp=1):
    """run <module>/<kernel>.comp of the reference on the CPU, one invocation per (x, y, z) < (wd, ht, dp).
    bindings: per descriptor binding of set 1 either (array, f16) or a list of those for an array connector; arrays are float32
    (h, w) or (h, w, c), written in place for image2D bindings (f16: stores round to half precision)."""
    flat, counts = [], []
    for b in bindings:
        items = b if isinstance(b, list) else [b]
        counts.append(len(items))
        for a, f16 in items:
            assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
            flat.append(ShaderImage(a.ctypes.data_as(C.POINTER(C.c_float)), a.shape[1], a.shape[0], 1 if a.ndim == 2 else a.shape[2], int(f16)))
    arr = (ShaderImage * len(flat))(*flat)
    fn = getattr(ref_shader_lib(), "shader_%s_%s" % (module.replace("-", "_"), kernel))
    r = fn(bytes(params), len(bytes(params)), bytes(push), len(bytes(push)), arr, (C.c_int * len(counts))(*counts), len(counts), wd, ht, dp)
    assert r == 0, "shader_%s_%s failed: %d" % (module, kernel, r)


def ref_pipeline_run(text, raw, trace=None):
    """execute the REFERENCE's own node graph with the REFERENCE's own shaders, on the CPU: `text` is what ref_graph_describe() wrote
    (module pass of the reference's graph code: nodes, dispatch sizes, push constants, connector formats, wiring, parameter and
    committed blocks), every node runs <name>/<kernel>.comp through ref_shader() on images laid out as its connectors say (f16
    connectors round on store).  raw: (h, w) uint16 mosaic, read by the first node's consumers as UNORM (x / 65535, allocate.h:72-81).
    the module chain is linear (default darkroom graphs): a module's `input` is the output of the module before it.
    returns the image the sink receives, (h, w, 4) float32.  trace: optional dict filled with {"module:kernel#node": outputs}."""
    mods, cur, node = [], None, None
    for ln in text.splitlines():
        t = ln.split()
        if ln.startswith("module "):
            cur = dict(name=t[1], params=b"", committed=None, mconn=[], nodes=[])
            mods.append(cur)
        elif ln.startswith(" params"):
            cur["params"] = bytes.fromhex(t[1]) if len(t) > 1 else b""
        elif ln.startswith(" committed"):
            cur["committed"] = bytes.fromhex(t[1])
        elif ln.startswith(" mconn"):
            name, typ, chan, fmt = t[2].split(":")
            assert t[-1] == "bypass=-1", "bypassed modules are not handled"
            cur["mconn"].append(dict(name=name, type=typ))
        elif ln.startswith(" node"):
            wd, ht, dp = [int(x) for x in t[3].split("x")]
            pc = t[4].split(":", 1)[1]
            node = dict(name=t[2], wd=wd, ht=ht, dp=dp, push=b"".join(int(x, 16).to_bytes(4, "little") for x in pc.split(",")) if pc else b"", conn=[], out=None)
            cur["nodes"].append(node)
        elif ln.startswith("  conn"):
            name, typ, chan, fmt = t[2].split(":")
            w, h = [int(x) for x in t[3].split("/")[1].split("x")]
            nchan = 1 if chan in ("rggb", "rgbx", "ssbo") or len(chan) == 1 else (2 if len(chan) == 2 else 4)
            node["conn"].append(dict(name=name, type=typ, chan=nchan, fmt=fmt, w=w, h=h, al=max(1, int(t[4][3:])), link=t[5], buf=None))
    h, w = raw.shape
    unorm = np.ascontiguousarray(raw.astype(np.float32) / np.float32(65535.0))

    def module_output(mi):
        """the node connector that carries module mi's output"""
        oc = [i for i, c in enumerate(mods[mi]["mconn"]) if c["type"] in ("write", "source") and c["name"] == "output"][0]
        for n in mods[mi]["nodes"]:
            for c in n["conn"]:
                if c["type"] in ("write", "source") and c["link"] == "mod.%d" % oc:
                    return mi, n, c
        raise AssertionError("module %s has no node on its output" % mods[mi]["name"])

    def run(mi, n):
        if n["out"] is not None:
            return
        n["out"] = True
        m = mods[mi]
        if n["name"] == "i-raw:main" or n["name"] == "i-mlv:main":
            n["conn"][0]["buf"] = [unorm]
            return
        if n["name"] == "i-lut:main":                # the table the module's filename parameter names (core/lut.h), as stored
            fn = m["params"].split(b"\0", 1)[0].decode()
            with open(fn, "rb") as f:
                magic, version, chan, dtype, lw, lh = struct.unpack("<IHBBII", f.read(16))
                a = np.frombuffer(f.read(), dtype=np.float16 if dtype == 0 else np.float32, count=lw * lh * chan)
            n["conn"][0]["buf"] = [np.ascontiguousarray(a.astype(np.float32).reshape(lh, lw, chan))]
            return
        binds, first_input = [], None
        for c in n["conn"]:
            if c["type"] in ("write",):
                c["buf"] = [np.zeros((c["h"], c["w"]) if c["chan"] == 1 else (c["h"], c["w"], c["chan"]), np.float32) for _ in range(c["al"])]
                binds.append([(b, 0 if c["fmt"] in ("f32", "ui8") else 1) for b in c["buf"]])   # ui8: stored as float, quantised below
                continue
            src = None
            if c["link"].startswith("n"):
                k, cc = [int(x) for x in c["link"][1:].split(".")]
                run(mi, m["nodes"][k])
                src = m["nodes"][k]["conn"][cc]["buf"]
            elif c["link"].startswith("mod.") and m["mconn"][int(c["link"][4:])]["name"] == "input":
                prev = mi - 1
                while mods[prev]["name"] == "i-lut":  # side inputs sit anywhere before their reader in the execution order
                    prev -= 1
                pm, pn, pc = module_output(prev)
                run(pm, pn)
                src = pc["buf"]
            elif c["link"].startswith("mod.") and m["mconn"][int(c["link"][4:])]["name"] in ("clut", "abney", "spectra"):
                # a lut connector that is wired: the i-lut module whose file carries the connector's name (the tests' convention)
                want = m["mconn"][int(c["link"][4:])]["name"].encode() + b".lut"
                lm = [k for k, x in enumerate(mods) if x["name"] == "i-lut" and want in x["params"].split(b"\0", 1)[0]]
                assert len(lm) == 1, (want, lm)
                pm, pn, pc = module_output(lm[0])
                run(pm, pn)
                src = pc["buf"]
            if src is None:                       # unconnected lut / gainmap inputs: the reference binds a dummy (graph-run-nodes-allocate.h)
                src = first_input
            assert src is not None, (m["name"], n["name"], c["name"])
            if first_input is None:
                first_input = src
            c["buf"] = src
            binds.append([(b, 0) for b in src])
        if n["name"].startswith("o-") or n["name"] == "display:main":
            return
        module, kernel = n["name"].split(":")
        ref_shader(module, kernel, m["committed"] if m["committed"] is not None else m["params"] + b"\0" * 16, n["push"] + b"\0" * 16,
                   [b if len(b) > 1 else b[0] for b in binds], n["wd"], n["ht"], n["dp"])
        for c in n["conn"]:
            if c["type"] == "write" and c["fmt"] == "ui8":
                # imageStore to an rgba8 UNORM image: clamp to [0, 1], scale, round to nearest even; read back as the 8 bit value
                for b in c["buf"]:
                    with np.errstate(invalid="ignore"):
                        b[...] = np.rint(np.clip(np.nan_to_num(b, nan=0.0), 0.0, 1.0) * np.float32(255.0))
        if trace is not None:
            trace["%s#%d" % (n["name"], m["nodes"].index(n))] = [c["buf"] for c in n["conn"] if c["type"] == "write"]

    sink = mods[-1]["nodes"][0]
    run(len(mods) - 1, sink)
    out = sink["conn"][0]["buf"][0]
    return out
