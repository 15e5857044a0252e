"""CPU: rtpp_b200/csrc/glibc_math.cuh — the restatement of glibc 2.39's sinf / cosf / atanf / atan2f that the device samplers use —
compiled for the HOST (oracle/librtpp_oracle.so) against this machine's own libm: same bits on every sampled argument.  (The
device instantiation is compared with libm over all 2^32 arguments in tests/test_glibc_math_gpu.py.)  The restatement was
first validated against libm.so.6's two ifunc builds of sinf / cosf (baseline and -mfma) over all 2^32 arguments; on a host
with another libm these tests say so."""
import numpy as np
import pytest


def test_sinf_cosf_atanf_checksums_match_libm(oracle):
    fma = oracle.glibc_host_fma()
    # 32 chunks of 2^22 arguments across the binary32 range (tiny, < pi/4, the reduction ranges, huge, negative, inf / nan) and the
    # neighbourhoods of the arguments on which the two libm builds differ
    firsts = [i << 27 for i in range(32)] + [0x3f490fdb - (1 << 21), 0x42f00000 - (1 << 21), 0x4255b0a9 - 100, 0x418a3adb - 100, 0x7f800000 - (1 << 21)]
    for fn in (0, 1, 2):
        for first in firsts:
            want = oracle.glibc_checksum(0, fn, first, 1 << 22, nthreads=8)
            got = oracle.glibc_checksum(1 + fma, fn, first, 1 << 22, nthreads=8)
            assert got == want, (fn, hex(first), fma)


def test_the_two_libm_builds_differ_only_where_expected(oracle):
    """The port's fused and unfused instantiations agree except on the known arguments (12 for sinf): the 2^22 neighbourhood of one
    of them must differ, a neighbourhood without any must not."""
    assert oracle.glibc_checksum(1, 0, 0x4255b0a9 - 100, 200, 1) != oracle.glibc_checksum(2, 0, 0x4255b0a9 - 100, 200, 1)
    assert oracle.glibc_checksum(1, 0, 0x3f800000, 1 << 20, 8) == oracle.glibc_checksum(2, 0, 0x3f800000, 1 << 20, 8)


def test_atan2f_matches_libm(oracle):
    rng = np.random.default_rng(4)
    yx = np.concatenate([
        rng.standard_normal((1 << 19, 2)).astype(np.float32) * np.float32(1000),
        (rng.random((1 << 19, 2), dtype=np.float32) - np.float32(0.5)) * np.float32(8192),  # goal - root differences on the maps
        (rng.integers(0, 1 << 32, (1 << 19, 2), dtype=np.uint64).astype(np.uint32)).view(np.float32),  # any bit patterns
        np.array([[0, 0], [0, -0.0], [-0.0, 0], [-0.0, -0.0], [1, 0], [0, 1], [-1, 0], [0, -1], [np.inf, 1], [1, np.inf], [np.inf, np.inf],
                  [-np.inf, -np.inf], [np.inf, -np.inf], [np.nan, 1], [1, np.nan], [1e-30, 1e30], [1e30, 1e-30], [1e30, -1e-30], [3, 1], [-3, 1.0],
                  [1, 1], [2.5, 1], [0.5, 1], [0.45, 1], [0.7, 1], [1.2, 1]], np.float32)])
    a, b = oracle.glibc_atan2f(0, yx), oracle.glibc_atan2f(1, yx)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), yx[~same][:5]
