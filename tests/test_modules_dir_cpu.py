"""module discovery (src/pipe/global.c:86-415): the library's built-in connector / parameter tables against the reference's own
modules/<name>/{connectors,params} files, and a vkdt checkout named with vkb_set_basedir() as the source of truth (every module
directory of the checkout then parses in a cfg).  the comparison with the files needs /root/reference; the rest runs anywhere."""
import os
import struct
import numpy as np
import pytest

from vkdt_b200 import api

REF = "/root/reference/src/pipe"
OURS = ["i-raw", "i-mlv", "i-pfm", "denoise", "hilite", "demosaic", "crop", "colour", "filmcurv", "llap", "grade", "colenc", "o-pfm", "o-jpg", "o-null",
        "display", "hist", "zones", "lens", "pick"]


def _tok(s):
    return s[:8]


def _file_tables(name):
    """parse the reference's files the way global.c does: tokens of <= 8 chars, defaults back to back, missing defaults 0."""
    d = os.path.join(REF, "modules", name)
    conns = []
    for ln in open(os.path.join(d, "connectors")).read().splitlines():
        if ln.strip():
            f = ln.split(":")
            conns.append("connector %s:%s:%s:%s" % tuple(_tok(x) for x in f[:4]))
    params, off = [], 0
    pf = os.path.join(d, "params")
    if os.path.exists(pf):
        for ln in open(pf).read().splitlines():
            if not ln.strip():
                continue
            f = ln.split(":")
            name_, typ, cnt = _tok(f[0]), _tok(f[1]), int(f[2])
            vals = f[3:]
            if typ == "float":
                blob = b"".join(struct.pack("<f", np.float32(float(vals[i])) if i < len(vals) and vals[i] != "" else 0.0) for i in range(cnt))
            elif typ == "int":
                blob = b"".join(struct.pack("<i", int(vals[i]) if i < len(vals) and vals[i] != "" else 0) for i in range(cnt))
            else:
                txt = ":".join(vals).encode()[:cnt - 1]
                blob = txt + b"\0" * (cnt - len(txt))
            params.append("param %s:%s:%d:%d:%s" % (name_, typ, cnt, off, blob.hex()))
            off += len(blob)
    return conns + params


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")
@pytest.mark.parametrize("name", OURS)
def test_builtin_tables_are_the_reference_files(name):
    api.set_basedir("")
    got = api.module_describe(name).splitlines()
    want = _file_tables(name)
    if name == "i-raw":
        want = [w.replace("connector output:source:*:ui16", "connector output:source:*:ui16") for w in want]
    assert got == want, "\n".join(["%s | %s" % (a, b) for a, b in zip(got, want) if a != b][:5])


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")
def test_basedir_makes_the_checkout_the_source_of_truth():
    try:
        api.set_basedir(REF)
        for name in OURS:
            assert api.module_describe(name).splitlines() == _file_tables(name), name
        # a module this library has no kernels for: its files register it, a cfg naming it parses
        assert api.module_describe("exposure").splitlines() == _file_tables("exposure")
        g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw") + "module:exposure:01\nmodule:grain:01\nparam:exposure:01:exposure:0.5\n")
        raw = np.zeros((384, 512), dtype=np.uint16)
        g.set_source(raw.ctypes.data, api.raw_params(512, 384))
        plan = g.plan()                               # unconnected extra modules do not reach the sink: the path plans as before
        assert "b200_llapfin" in plan and "exposure" not in plan
        g.close()
        # ... and connected into the path it fails loudly: no kernel for it
        cfg = api.DARKROOM_CFG.format(src="i-raw").replace("connect:colour:01:output:filmcurv:01:input\n",
            "module:exposure:01\nconnect:colour:01:output:exposure:01:input\nconnect:exposure:01:output:filmcurv:01:input\n")
        g = api.Graph(cfg_text=cfg)
        g.set_source(raw.ctypes.data, api.raw_params(512, 384))
        with pytest.raises(api.VkbError) as e:
            g.plan()
        assert "exposure" in str(e.value)
        g.close()
    finally:
        api.set_basedir("")


def test_unknown_module_without_basedir_is_a_config_warning():
    api.set_basedir("")
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    assert g.line("module:exposure:01") > 0          # graph-io.c: unknown module, warning, the reader goes on
    g.close()
