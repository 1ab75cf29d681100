"""host-side file parsers (dng, lossless jpeg, mlv container, pfm header) against truncated and corrupted input:
every outcome but a crash or a hang is fine (an error code, or a decode of whatever the bytes say)."""
import numpy as np
import pytest

from vkdt_b200 import api, synth


def _mutations(data, rng, n):
    data = bytearray(data)
    yield bytes(data[:0])
    for _ in range(n):
        d = bytearray(data)
        kind = rng.integers(0, 4)
        if kind == 0:
            d = d[:rng.integers(0, len(d))]                                  # truncate
        elif kind == 1:
            for _ in range(rng.integers(1, 8)):
                d[rng.integers(0, len(d))] = rng.integers(0, 256)             # flip bytes anywhere
        elif kind == 2:
            for _ in range(rng.integers(1, 8)):
                d[rng.integers(0, min(len(d), 256))] = rng.integers(0, 256)   # flip bytes in the headers
        else:
            i = rng.integers(0, max(1, len(d) - 4))
            d[i:i + 4] = b"\xff\xff\xff\xff"                                  # huge counts / offsets
        yield bytes(d)


def test_dng_reader_survives_corruption(tmp_path):
    rng = np.random.default_rng(1)
    fn = str(tmp_path / "a.dng")
    synth.write_dng(fn, rng.integers(0, 65535, (24, 36), dtype=np.uint16), cfa=((1, 2), (0, 1)), active_area=(2, 2, 22, 34))
    good = open(fn, "rb").read()
    bad = str(tmp_path / "b.dng")
    ok = 0
    for m in _mutations(good, rng, 400):
        open(bad, "wb").write(m)
        try:
            p, ox, oy = api.dng_info(bad)
            ok += 1
            assert p.width <= 36 + 2 or p.width * p.height <= (1 << 31)
        except api.VkbError:
            pass
    assert ok > 0      # some mutations only touch pixels


def test_lj92_decoder_survives_corruption():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 1 << 14, (16, 24)).astype(np.uint16)
    for comps in (1, 2):
        good = synth.lj92_encode(img, 14, comps, 1 if comps == 2 else 6)
        for m in _mutations(good, rng, 400):
            try:
                out, bits = api.lj92_decode(m)
                assert out.size <= 1 << 26
            except (api.VkbError, MemoryError, ValueError):
                pass


def test_mlv_and_pfm_sources_survive_corruption(tmp_path):
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:34, 0:64]
    pix = (2048 + 37 * xx + 11 * yy).astype(np.uint16)
    for lossless in (False, True):
        fn = str(tmp_path / "g.mlv")
        synth.write_mlv(fn, [pix, pix], bpp=14, lossless=lossless, camera_name="Canon EOS")
        good = open(fn, "rb").read()
        bad = str(tmp_path / "b.mlv")
        for m in _mutations(good, rng, 150):
            open(bad, "wb").write(m)
            g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-mlv"))
            g.line("param:i-mlv:main:filename:%s" % bad)
            try:
                g.plan()
            except api.VkbError:
                pass
            g.close()
    fn = str(tmp_path / "x.pfm")
    synth.write_pfm(fn, rng.random((12, 20, 3), dtype=np.float32))
    good = open(fn, "rb").read()
    bad = str(tmp_path / "b.pfm")
    for m in _mutations(good, rng, 100):
        open(bad, "wb").write(m)
        g = api.Graph(cfg_text="module:i-pfm:main\nmodule:filmcurv:01\nmodule:o-null:main\nconnect:i-pfm:main:output:filmcurv:01:input\n"
                               "connect:filmcurv:01:output:o-null:main:input\nparam:i-pfm:main:filename:%s\n" % bad, sink=None)
        try:
            g.plan()
        except api.VkbError:
            pass
        g.close()


def test_parsers_under_address_and_ub_sanitizer(tmp_path):
    """the same parsers compiled stand-alone with -fsanitize=address,undefined (tests/tools/parser_harness.cpp) and fed
    mutated files: out-of-bounds reads that happen to stay inside the process would pass the tests above."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "parser_harness")
    src = [os.path.join(root, "tests/tools/parser_harness.cpp")] + [os.path.join(root, "vkdt_b200/csrc/pipe", f) for f in ("dng.cpp", "lj92.cpp", "mlv.cpp")]
    r = subprocess.run([gxx, "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe] + src, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available: " + r.stderr[-200:])
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:34, 0:64]
    pix = (2048 + 37 * xx + 11 * yy).astype(np.uint16)
    seeds = {}
    fn = str(tmp_path / "a.dng")
    synth.write_dng(fn, rng.integers(0, 65535, (24, 36), dtype=np.uint16), cfa=((1, 2), (0, 1)), active_area=(2, 2, 22, 34))
    seeds["dng"] = [open(fn, "rb").read()]
    img = rng.integers(0, 1 << 14, (16, 24)).astype(np.uint16)
    seeds["lj92"] = [synth.lj92_encode(img, 14, 2, 1), synth.lj92_encode(img, 12, 1, 6)]
    seeds["mlv"] = []
    for lossless in (False, True):
        synth.write_mlv(fn, [pix, pix], bpp=14, lossless=lossless, camera_name="Canon EOS")
        seeds["mlv"].append(open(fn, "rb").read())
    for kind, goods in seeds.items():
        files = []
        for good in goods:
            for m in _mutations(good, rng, 250):
                f2 = str(tmp_path / ("%s_%d" % (kind, len(files))))
                open(f2, "wb").write(m)
                files.append(f2)
        for k in range(0, len(files), 64):
            r = subprocess.run([exe, kind] + files[k:k + 64], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-2000:]


def test_config_grammar_survives_garbage():
    """graph-io style lines with mangled tokens, counts and numbers: warnings or errors, never a crash; the graph still plans."""
    rng = np.random.default_rng(9)
    base = api.DARKROOM_CFG.format(src="i-raw").splitlines() + [
        "param:crop:01:crop:0.1:0.9:0.1:0.9", "paramsub:crop:01:crop:1:2:0.5", "paraminc:colour:01:exposure:0:1:0.25",
        "param:colour:01:rbmap:" + ":".join(["0.5"] * 144), "param:i-raw:main:filename:/nonexistent/file.dng", "frames:10", "fps:24"]
    alphabet = list(b":0123456789.-+eE abcxyz\t#%\xff\x00")
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    for _ in range(4000):
        line = bytearray(base[rng.integers(0, len(base))].encode())
        for _ in range(rng.integers(1, 6)):
            k = rng.integers(0, 4)
            if k == 0 and line:
                line[rng.integers(0, len(line))] = alphabet[rng.integers(0, len(alphabet))]
            elif k == 1:
                i = rng.integers(0, len(line) + 1)
                line[i:i] = bytes(rng.choice(alphabet, rng.integers(1, 40)).tolist())
            elif k == 2 and line:
                del line[rng.integers(0, len(line)):]
            else:
                line += b":" + str(rng.integers(-2**40, 2**40)).encode()
        api.lib.vkb_graph_read_config_line(g.h, bytes(line).replace(b"\x00", b"?"))
    raw = np.zeros((64, 96), np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(96, 64))
    try:                      # whatever the mangled lines wired up (self loops, cycles, dangling inputs): an error or a plan
        g.plan()
    except api.VkbError:
        pass
    g2 = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    g2.set_source(raw.ctypes.data, api.raw_params(96, 64))
    assert "sink o-pfm" in g2.plan()
