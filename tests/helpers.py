"""shared helpers of the parity tests: device upload/download in the engine's HBM layouts and f16 ulp metrics."""
import struct
import numpy as np


def f16_ulp_diff(a, b):
    """ulp distance between two arrays after rounding both to binary16 (nan == nan, +0 == -0)."""
    ha = np.asarray(a, dtype=np.float32).astype(np.float16).view(np.int16).astype(np.int32)
    hb = np.asarray(b, dtype=np.float32).astype(np.float16).view(np.int16).astype(np.int32)
    ka = np.where(ha < 0, -(ha & 0x7fff), ha)
    kb = np.where(hb < 0, -(hb & 0x7fff), hb)
    d = np.abs(ka - kb)
    nan = np.isnan(np.asarray(a, dtype=np.float32)) & np.isnan(np.asarray(b, dtype=np.float32))
    d[nan] = 0
    return d


def census(what, d, got=None, want=None):
    """VKB_CENSUS=<file>: append one line per comparison (max f16 ulp, share of bit-identical values, max abs) - the per
    kernel record of how far the strict kernels are from the restatement, kept under profiles/."""
    import os
    path = os.environ.get("VKB_CENSUS")
    if not path:
        return
    mabs = float(np.nanmax(np.abs(np.asarray(got, dtype=np.float64) - np.asarray(want, dtype=np.float64)))) if got is not None else -1.0
    with open(path, "a") as f:
        f.write("%-44s max_ulp %3d  identical %.6f  differing %d of %d  max_abs %.3g\n" % (what, int(d.max()), float((d == 0).mean()), int((d != 0).sum()), d.size, mabs))


def assert_f16_close(got, want, max_ulp=2, min_exact=0.98, what=""):
    d = f16_ulp_diff(got, want)
    census(what, d, got, want)
    exact = float((d == 0).mean())
    assert d.max() <= max_ulp and exact >= min_exact, "%s: max ulp %d (allowed %d), bit-identical %.4f (needed %.4f)" % (
        what, int(d.max()), max_ulp, exact, min_exact)
    return d


def assert_close_mixed(got, want, ulps=3, atol=4e-6, min_exact=0.85, what="", max_outliers=0.0, hard_atol=1e-3):
    """f16-ulp tolerance with an absolute floor: near black 1 - exp(-x) loses relative (not absolute) accuracy
    in any fp32 implementation, the reference's included."""
    got = np.asarray(got, dtype=np.float32).copy(); want = np.asarray(want, dtype=np.float32).copy()
    both_nan = np.isnan(got) & np.isnan(want)      # a NaN both sides produce (e.g. dead cov outputs) is agreement
    got[both_nan] = 0.0; want[both_nan] = 0.0
    assert not (np.isnan(got) | np.isnan(want)).any(), "%s: NaN on one side only" % what
    d = f16_ulp_diff(got, want)
    census(what, d, got, want)
    bad = (d > ulps) & (np.abs(got - want) > atol)
    exact = float((d == 0).mean())
    assert float(np.abs(got - want).max()) <= hard_atol, "%s: max abs %.3g" % (what, float(np.abs(got - want).max()))
    assert bad.mean() <= max_outliers and exact >= min_exact, "%s: %d outliers, max ulp %d, max abs %.3g, bit-identical %.4f" % (
        what, int(bad.sum()), int(d.max()), float(np.abs(got - want).max()), exact)


def parity_gate(got, want, what="", max_abs=1e-3, min_psnr=60.0):
    """BASELINE.json north_star, literally: float outputs within max-abs 1e-3 and PSNR >= 60 dB in linear rec2020.
    (the strict kernels are bit compatible with the restatement: what is measured here is normally 0 and 99 dB; the census
    file records it.)"""
    got = np.asarray(got, dtype=np.float32); want = np.asarray(want, dtype=np.float32)
    assert got.shape == want.shape, (got.shape, want.shape)
    both_nan = np.isnan(got) & np.isnan(want)
    assert not (np.isnan(got) ^ np.isnan(want)).any(), "%s: NaN on one side only" % what
    err = np.abs(np.where(both_nan, 0.0, got) - np.where(both_nan, 0.0, want))
    p = psnr(np.where(both_nan, 0.0, got), np.where(both_nan, 0.0, want))
    import os
    path = os.environ.get("VKB_CENSUS")
    if path:
        with open(path, "a") as f:
            f.write("%-44s graph: max_abs %.3g  psnr %.1f dB  differing %d of %d  > 1e-3: %d\n" % (what, float(err.max()), p, int((err != 0).sum()), err.size, int((err > 1e-3).sum())))
    assert p >= min_psnr, "%s: psnr %.1f dB < %.1f" % (what, p, min_psnr)
    assert float(err.max()) <= max_abs, "%s: max abs %.3g > %.3g (%d of %d values beyond it)" % (what, float(err.max()), max_abs, int((err > max_abs).sum()), err.size)
    return float(err.max()), p


def census_graph(what, got, want):
    import os
    path = os.environ.get("VKB_CENSUS")
    if path:
        err = np.abs(np.asarray(got, dtype=np.float64) - np.asarray(want, dtype=np.float64))
        with open(path, "a") as f:
            f.write("%-44s graph: max_abs %.3g  differing %d of %d  > 1e-3: %d\n" % (what, float(np.nanmax(err)), int((err != 0).sum()), err.size, int((err > 1e-3).sum())))


def psnr(got, want, peak=1.0):
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    mse = np.mean((got - want) ** 2)
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


# ---- device transfer (torch is plumbing only: device memory + streams) ----
def to_dev_f16(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).astype(np.float16)).cuda()


def to_dev_u16(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint16).view(np.int16)).cuda()


def dev_f16(h, w, c=1, layers=1):
    import torch
    shape = (layers, h, w, c) if layers > 1 else ((h, w, c) if c > 1 else (h, w))
    return torch.zeros(shape, dtype=torch.float16, device="cuda")


def dev_f32(h, w, c=4):
    import torch
    return torch.zeros((h, w, c), dtype=torch.float32, device="cuda")


def to_host(t):
    return t.float().cpu().numpy()


def fbits(*vals):
    return struct.pack("<%df" % len(vals), *vals)


def ibits(*vals):
    return struct.pack("<%di" % len(vals), *vals)


def ubits(*vals):
    return struct.pack("<%dI" % len(vals), *vals)


def levels(w, h, maxl):
    """pyramid level sizes as built by hilite/main.c:44-88 and llap/main.c:39-88: returns list of (w,h), index 0 = finest."""
    out = [(w, h)]
    cw, ch = (w - 1) // 2 + 1, (h - 1) // 2 + 1
    for l in range(1, maxl):
        out.append((cw, ch))
        cw, ch = (cw - 1) // 2 + 1, (ch - 1) // 2 + 1
        if cw <= 1 or ch <= 1 or l + 1 == maxl:
            break
    return out
