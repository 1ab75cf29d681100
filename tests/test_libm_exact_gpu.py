"""the strict kernels' exp / exp2 / log2 / pow ON THE DEVICE against libm (through the oracle's o_libm_apply): bit for bit,
special values included.  the fast build's SFU versions stay within a few ulp on the ranges the path uses."""
import ctypes as C
import numpy as np
import pytest
from helpers import ubits

pytestmark = pytest.mark.gpu


def _device(gpu, op, a, b, mode):
    import torch
    gpu.set_mode(mode)
    try:
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        do = torch.zeros_like(da)
        n = a.size
        w = 1 << 16
        assert n % w == 0
        gpu.dispatch("b200", "libm", [gpu.image(da, w, n // w, 1, "f32"), gpu.image(db, w, n // w, 1, "f32"), gpu.image(do, w, n // w, 1, "f32")],
                     push=ubits(op))
        torch.cuda.synchronize()
        return do.cpu().numpy()
    finally:
        gpu.set_mode(gpu.MODE_STRICT)


def _libm(oracle, op, a, b):
    out = np.zeros_like(a)
    fp = lambda x: x.ctypes.data_as(C.c_void_p)
    oracle.lib().o_libm_apply.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    oracle.lib().o_libm_apply(op, a.size, fp(a), fp(b), fp(out))
    return out


def _same(x, y):
    return (x.view(np.uint32) == y.view(np.uint32)) | (np.isnan(x) & np.isnan(y))


@pytest.mark.parametrize("op", [0, 1, 2, 3])
def test_strict_transcendentals_are_libm_bit_for_bit(gpu, oracle, op):
    rng = np.random.default_rng(1234 + op)
    n = 1 << 24
    a = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)     # every kind of bit pattern
    b = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    # the ranges the path lives in, densely: arguments in (0, 8), exponents in (-8, 8)
    a[n // 2:] = (rng.random(n // 2, dtype=np.float32) * 8.0 - (4.0 if op < 2 else 0.0)).astype(np.float32)
    b[n // 2:] = (rng.random(n // 2, dtype=np.float32) * 16.0 - 8.0).astype(np.float32)
    specials = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1.17549435e-38, 3.4028235e38, 88.72284, -103.97, 127.99999, -149.5,
                         0.5, 2.0, 3.0, 4.0, 16.0, 0.8, 1.0 / 3.0], dtype=np.float32)
    k = specials.size
    a[:k * k] = np.repeat(specials, k)
    b[:k * k] = np.tile(specials, k)
    got = _device(gpu, op, a, b, gpu.MODE_STRICT)
    want = _libm(oracle, op, a, b)
    bad = ~_same(got, want)
    assert not bad.any(), "op %d: %d of %d differ, e.g. f(%r, %r) = %r, libm %r" % (op, int(bad.sum()), n, a[bad][0], b[bad][0], got[bad][0], want[bad][0])


def test_fast_transcendentals_stay_close(gpu, oracle):
    rng = np.random.default_rng(7)
    n = 1 << 22
    a = (rng.random(n, dtype=np.float32) * 4.0 + 1e-3).astype(np.float32)
    b = (rng.random(n, dtype=np.float32) * 3.0).astype(np.float32)
    for op, arg in ((0, -a), (3, a)):
        got = _device(gpu, op, arg, b, gpu.MODE_FAST)
        want = _libm(oracle, op, arg, b)
        rel = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)
        assert rel.max() < 4e-6, (op, float(rel.max()))
