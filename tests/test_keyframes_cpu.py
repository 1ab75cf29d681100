"""keyframes and feedback lines: `keyframe:frame:module:inst:param:beg:end:values` in its five spellings
(graph-io.c:139-153, :243-247) and dt_graph_apply_keyframes (graph.c:1025-1100, easings of anim.h:13-47), against a restatement
of those lines written here in numpy.  host only; tests/test_graph_gpu.py develops a keyframed sequence."""
import struct
import numpy as np
import pytest

from vkdt_b200 import api

EASE = {"keyframe": lambda t: t, "keyFRAME": lambda t: 1.0 if t > 0.5 else 0.0, "Keyframe": lambda t: t * t * t,
        "keyframE": lambda t: 1.0 - (1.0 - t) ** 3, "KeyframE": lambda t: 3.0 * t * t - 2.0 * t * t * t}


def _param(g, module, inst, off, n):
    for ln in g.state().splitlines():
        if ln.startswith("%s:%s " % (module, inst)):
            blob = bytes.fromhex(ln.split()[1])
            return np.frombuffer(blob[off:off + 4 * n], dtype=np.float32).copy()
    raise KeyError(module)


def _graph(extra):
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw") + "frames:40\n", sink=None)
    for ln in extra:
        assert g.line(ln) == 0, ln
    return g


@pytest.mark.parametrize("cmd", sorted(EASE))
def test_float_keyframes_interpolate_with_their_easing(cmd):
    # colour:exposure is the first parameter (offset 0); llap: sigma, shadows, hilights, clarity
    g = _graph(["keyframe:5:colour:01:exposure:0:1:-1.0", "%s:15:colour:01:exposure:0:1:2.0" % cmd, "keyframe:30:colour:01:exposure:0:1:0.5"])
    for f in (0, 5, 7, 10, 14, 15, 20, 29, 30, 39):
        g.set_frame(f)
        g.apply_keyframes()
        got = float(_param(g, "colour", "01", 0, 1)[0])
        if f < 5:      # all keyframes ahead: the earliest is both "current" and "next" (graph.c:1039-1053), it is applied as is
            want = -1.0
        elif f < 15:
            t = np.float32(np.float32(f - 5) / np.float32(10)); e = np.float32(EASE[cmd](float(t))); want = e * np.float32(2.0) + (np.float32(1.0) - e) * np.float32(-1.0)
        elif f < 30:   # towards frame 30, whose line is a plain `keyframe`: linear
            t = np.float32(np.float32(f - 15) / np.float32(15)); want = t * np.float32(0.5) + (np.float32(1.0) - t) * np.float32(2.0)
        else:
            want = 0.5
        assert abs(got - float(want)) <= 1e-6 * max(1.0, abs(float(want))), (cmd, f, got, float(want))
    g.close()


def test_keyframes_leave_other_parameters_alone_and_ints_step():
    g = _graph(["param:llap:01:clarity:0.33", "keyframe:0:llap:01:sigma:0:1:0.1", "keyframe:10:llap:01:sigma:0:1:0.3",
                "keyframe:0:filmcurv:01:colour:0:1:1", "keyframe:10:filmcurv:01:colour:0:1:3"])
    for f, sig, col in ((0, 0.1, 1), (5, 0.2, 1), (9, 0.28, 1), (10, 0.3, 3), (12, 0.3, 3)):
        g.set_frame(f)
        g.apply_keyframes()
        ll = _param(g, "llap", "01", 0, 4)
        assert abs(ll[0] - sig) < 1e-6 and abs(ll[3] - 0.33) < 1e-7 and abs(ll[1] - 1.0) < 1e-7
        for ln in g.state().splitlines():
            if ln.startswith("filmcurv:01 "):
                assert struct.unpack("<i", bytes.fromhex(ln.split()[1])[12:16])[0] == col   # int parameters are copied, never blended (graph.c:1094-1097)
    g.close()


def test_a_keyframe_line_replaces_the_one_of_the_same_frame():
    g = _graph(["keyframe:4:colour:01:exposure:0:1:1.0", "keyframe:4:colour:01:exposure:0:1:3.0"])
    g.set_frame(4)
    g.apply_keyframes()
    assert abs(float(_param(g, "colour", "01", 0, 1)[0]) - 3.0) < 1e-7
    g.close()


def test_feedback_lines_flag_the_graph():
    g = _graph(["module:grade:02"])
    assert not g.has_feedback()
    assert g.line("feedback:grade:01:output:grade:02:input") == 0
    assert g.has_feedback()
    g.close()
    # a graph that reaches a feedback connector plans with a double buffered owner (tests/test_feedback.py runs one)
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    assert g.line("feedback:grade:01:output:colour:01:spectra") == 0
    raw = np.zeros((384, 512), dtype=np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(512, 384))
    assert "[x2]" in g.plan()
    g.close()
