"""band split on the device: one frame developed by several device slots (vkb_graph_set_bands) has to be BIT FOR BIT the
frame one GPU develops (BASELINE.json config 5).  slots sharing GPU 0 exercise the whole mechanism
(band limited launches, halo pulls between pools, the all-gather in front of the small pyramid levels, event ordering) on a
one GPU box; with two or more GPUs the same test runs over NVLink peer access."""
import numpy as np
import pytest

from vkdt_b200 import synth

pytestmark = pytest.mark.gpu

WB = (2.0, 1.0, 1.5)
CAM = (0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8)


def _develop(gpu, raw, devices=None, mode=None, frames=1):
    h, w = raw.shape
    g = gpu.Graph(cfg_text=gpu.DARKROOM_CFG.format(src="i-raw"))
    assert g.line("param:denoise:01:strength:0.4") == 0
    if mode is not None:
        g.set_mode(mode)
    buf = np.ascontiguousarray(raw)
    g.set_source(buf.ctypes.data, gpu.raw_params(w, h, wb=WB, cam_to_rec2020=CAM, noise_a=100.0, noise_b=2.0))
    g.set_sink_layout(gpu.SINK_RGB_F32)
    if devices:
        g.set_bands(devices)
    g.set_sink_buffer(None, 0)
    g.run()
    ow, oh = g.sink_size()
    out = np.zeros((oh, ow, 3), dtype=np.float32)
    g.set_sink_buffer(out.ctypes.data, out.nbytes)
    for _ in range(frames):
        out[:] = -1.0
        g.run(gpu.RUN_RECORD | gpu.RUN_UPLOAD | gpu.RUN_DOWNLOAD | gpu.RUN_WAIT)
    stats = g.band_stats() if devices else None
    g.close()
    return out, stats


@pytest.mark.parametrize("dims,n", [((1536, 1024), 2), ((2000, 1234), 3), ((3072, 2304), 4), ((1024, 2560), 8)])
def test_bands_on_one_gpu_are_bit_identical(gpu, dims, n):
    w, h = dims
    raw = synth.mosaic(w, h, seed=211)
    one, _ = _develop(gpu, raw)
    banded, stats = _develop(gpu, raw, devices=[0] * n, frames=3)      # three frames: the pools are recycled between frames
    assert banded.shape == one.shape
    diff = banded != one
    print("%dx%d on %d slots: %d pulls, %.2f MB pulled per frame (%.2f %% of the output), %d kernel launches" % (
        w, h, n, stats["pulls"], stats["bytes_total"] / 1e6, 100.0 * stats["bytes_total"] / one.nbytes, stats["launches"]))
    assert not diff.any(), "%d of %d values differ, first at %s" % (int(diff.sum()), diff.size, np.argwhere(diff)[0])


def test_bands_fast_mode_bit_identical(gpu):
    raw = synth.mosaic(1536, 1024, seed=5)
    one, _ = _develop(gpu, raw, mode=gpu.MODE_FAST)
    banded, _ = _develop(gpu, raw, devices=[0, 0, 0], mode=gpu.MODE_FAST)
    assert np.array_equal(one, banded)


def test_bands_across_gpus(gpu):
    """real peer access: needs two GPUs (the 8 GPU scaling run of bench.py --workload still201 covers the rest)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU")
    raw = synth.mosaic(3072, 2304, seed=77)
    one, _ = _develop(gpu, raw)
    banded, stats = _develop(gpu, raw, devices=list(range(min(n, 8))), frames=3)
    print("%d GPUs: %.2f MB over NVLink per frame" % (min(n, 8), stats["bytes_total"] / 1e6))
    assert np.array_equal(one, banded)
