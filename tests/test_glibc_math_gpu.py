"""GPU: the device instantiation of csrc/glibc_math.cuh (sinf, cosf, atanf: ALL 2^32 arguments, by checksums of 2^24-argument
chunks; atan2f: random, map-domain and special argument pairs) against the libm of the host the test runs on — the device
samplers (informed ellipse, RTRRT.cpp:241-262 / AMRRT.cpp:415-432) must return glibc's bits."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fn,name", [(0, "sinf"), (1, "cosf"), (2, "atanf")])
def test_device_matches_host_libm_on_every_binary32(ctx, oracle, fn, name):
    fma = ctx.L.rtpp_glibc_host_fma()
    assert fma == oracle.glibc_host_fma()
    bad = []
    for chunk in range(256):
        first, count = chunk << 24, 1 << 24
        out = C.c_ulonglong()
        ctx.call("glibc_checksum", fn, fma, C.c_uint32(first), C.c_ulonglong(count), C.byref(out))
        if out.value != oracle.glibc_checksum(0, fn, first, count, nthreads=16):
            bad.append(hex(first))
    assert not bad, (name, bad[:8])


def test_device_atan2f_matches_host_libm(ctx, oracle):
    rng = np.random.default_rng(5)
    yx = np.concatenate([
        rng.standard_normal((1 << 22, 2)).astype(np.float32) * np.float32(1000),
        (rng.random((1 << 22, 2), dtype=np.float32) - np.float32(0.5)) * np.float32(8192),
        (rng.integers(0, 1 << 32, (1 << 22, 2), dtype=np.uint64).astype(np.uint32)).view(np.float32),
        np.array([[0, 0], [0, -0.0], [-0.0, 0], [-0.0, -0.0], [1, 0], [0, 1], [-1, 0], [0, -1], [np.inf, 1], [1, np.inf], [np.inf, np.inf],
                  [-np.inf, -np.inf], [np.inf, -np.inf], [np.nan, 1], [1, np.nan], [1e-30, 1e30], [1e30, 1e-30], [1e30, -1e-30]], np.float32)])
    yx = np.ascontiguousarray(yx, np.float32)
    got = np.zeros(len(yx), np.float32)
    ctx.call("glibc_eval", 3, 0, yx.ctypes.data_as(C.c_void_p), C.c_size_t(len(yx)), got.ctypes.data_as(C.c_void_p))
    want = oracle.glibc_atan2f(0, yx)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (yx[~same][:5], got[~same][:5], want[~same][:5])
