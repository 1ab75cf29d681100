"""band split, host half (vkb_graph_band_plan, no GPU): what every device slot computes, pulls and waits for per launch.
checked here: the bands of a launch tile its image exactly (or every slot computes a small level whole), a slot only pulls
rows it does not compute itself, from a slot that computed them in an earlier launch, and waits for that slot; uploads cover
what the first launches read; the sink's rows are downloaded exactly once."""
import re
import numpy as np
import pytest

from vkdt_b200 import api


def _plan(w, h, n, strength=0.4):
    g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
    if strength > 0:
        assert g.line("param:denoise:01:strength:%g" % strength) == 0
    raw = np.zeros((h, w), dtype=np.uint16)
    g.set_source(raw.ctypes.data, api.raw_params(w, h, noise_a=100.0, noise_b=2.0))
    g.set_sink_layout(api.SINK_RGB_F32)
    g.set_bands([0] * n)
    text = g.band_plan()
    g.close()
    return text


def _parse(text):
    steps, down, up = {}, [], []
    for ln in text.splitlines():
        m = re.match(r"launch\s+(\d+) dev (\d+) (\S+) rows(.*)", ln)
        if m:
            li, d, name, rest = int(m.group(1)), int(m.group(2)), m.group(3), m.group(4)
            comp = [(int(a), int(b)) for a, b in re.findall(r" \[(\d+),(\d+)\)", rest.split(" pull")[0].split(" wait")[0])]
            pulls = [(int(b), int(a), int(c), int(s)) for b, a, c, s in re.findall(r"pull b(\d+)\[(\d+),(\d+)\)<-(\d+)", rest)]
            waits = [int(x) for x in re.findall(r"wait (\d+)", rest)]
            steps.setdefault(li, {})[d] = (name, comp, pulls, waits)
        m = re.match(r"download dev (\d+) b(\d+)\[(\d+),(\d+)\)", ln)
        if m:
            down.append(tuple(int(x) for x in m.groups()))
        m = re.match(r"upload dev (\d+) b(\d+)\[(\d+),(\d+)\)", ln)
        if m:
            up.append(tuple(int(x) for x in m.groups()))
    return steps, down, up


@pytest.mark.parametrize("w,h,n", [(1536, 1024, 2), (2000, 1234, 3), (4096, 3072, 8), (1024, 4000, 4)])
def test_band_plan_tiles_every_launch(w, h, n):
    steps, down, up = _parse(_plan(w, h, n))
    assert len(steps) >= 40
    for li, per_dev in steps.items():
        assert sorted(per_dev) == list(range(n))
        rows = [sum(b - a for a, b in per_dev[d][1]) for d in range(n)]
        sets = [set(r for a, b in per_dev[d][1] for r in range(a, b)) for d in range(n)]
        union = set().union(*sets)
        total = max(union) + 1
        assert union == set(range(total)), (li, per_dev[0][0])                       # no row left out
        if sum(rows) != total:                                                        # not a partition: a small level, whole on every slot
            assert all(s == union for s in sets) and total < 32 * n + 64, (li, per_dev[0][0], rows, total)
        for d in range(n):
            name, comp, pulls, waits = per_dev[d]
            for buf, a, b, src in pulls:
                assert src != d and src in waits and b > a
    # the sink: every output row downloaded by exactly one slot
    cover = sorted((a, b) for _, _, a, b in down)
    assert cover[0][0] == 0 and all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1)) and cover[-1][1] == h - 6
    # the source: the slots' uploads cover the mosaic (with overlap: halos of the first stencils)
    rows = set(r for _, _, a, b in up for r in range(a, b))
    assert rows == set(range(h))


def test_band_plan_needs_the_banded_kernels():
    """denoise off runs (denoise, noop), which has no band description: the plan refuses instead of guessing."""
    with pytest.raises(api.VkbError):
        _plan(1536, 1024, 2, strength=0.0)


def test_band_pull_volume_is_small():
    """the halo traffic of a 12 MP frame on 4 slots stays below 3 % of the frame's pooled bytes."""
    text = _plan(4096, 3072, 4)
    pulled = sum(int(x) for x in re.findall(r"pulls (\d+) bytes", text))
    assert 0 < pulled < 0.03 * 4096 * 3072 * 40
