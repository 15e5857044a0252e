// glibc_math.cuh — sinf / cosf / atanf / atan2f with the bits of GNU libc 2.39 on x86-64, for host AND device.
//
// The planners' informed-ellipse samplers (reference RTRRT.cpp:241-262, AMRRT.cpp:415-432) and DrawDObsmap (RTRRT.cpp:145-152)
// call glibc's binary32 transcendentals; their results decide which pixel a sample falls on, so a device sampler has to
// return the SAME bits.  None of these routines is correctly rounded; they are deterministic sequences of IEEE operations:
//   * sinf / cosf (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h): a binary64 argument reduction (|x| < 120: one
//     multiply, one truncation, one multiply-subtract; larger: 3 x 32-bit words of 4/pi) and binary64 polynomials.  libm picks
//     at load time between the baseline build and the -mfma build (ifunc: FMA && AVX2); in the latter EVERY a*b+c of the source
//     is one fused operation (read off the disassembly of libm.so.6), which changes the binary32 result for 12 of the 2^32
//     arguments of sinf.  kFma selects the build; rtpp_glibc_has_fma() tells which one the running host uses.
//   * atanf (sysdeps/ieee754/flt-32/s_atanf.c) and __ieee754_atan2f (e_atan2f.c): fdlibm's binary32 code, built for baseline
//     x86-64 (no ifunc): every operation rounds once.
// Validated on the host against libm's own code over all 2^32 arguments (sinf, cosf: both builds; atanf) and 2^32 random and
// structured argument pairs (atan2f) before it was used: tests/test_glibc_math.py; on the device: tests/test_glibc_math_gpu.py.
// Constants are the ones in libm.so.6's read-only data.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define GM_FN __device__ __forceinline__
#define GM_DMUL(a, b) __dmul_rn(a, b)
#define GM_DADD(a, b) __dadd_rn(a, b)
#define GM_DFMA(a, b, c) __fma_rn(a, b, c)
#define GM_FMUL(a, b) __fmul_rn(a, b)
#define GM_FADD(a, b) __fadd_rn(a, b)
#define GM_FSUB(a, b) __fsub_rn(a, b)
#define GM_FDIV(a, b) __fdiv_rn(a, b)
#define GM_F2D(a) ((double)(a))
#define GM_D2F(a) __double2float_rn(a)
#define GM_D2I(a) __double2int_rz(a)
#define GM_I2D(a) ((double)(a))
#define GM_L2D(a) __ll2double_rn(a)
#define GM_FBITS(a) __float_as_uint(a)
#define GM_BITSF(a) __uint_as_float(a)
#else
#if defined(__CUDACC__)
#define GM_FN __host__ __device__ inline
#else
#define GM_FN static inline
#endif
// host build: compiled with -ffp-contract=off (rtpp_b200/planners/Makefile, oracle/Makefile), so a * b + c below is two operations
#define GM_DMUL(a, b) ((a) * (b))
#define GM_DADD(a, b) ((a) + (b))
#define GM_DFMA(a, b, c) __builtin_fma(a, b, c)
#define GM_FMUL(a, b) ((float)((a) * (b)))
#define GM_FADD(a, b) ((float)((a) + (b)))
#define GM_FSUB(a, b) ((float)((a) - (b)))
#define GM_FDIV(a, b) ((float)((a) / (b)))
#define GM_F2D(a) ((double)(a))
#define GM_D2F(a) ((float)(a))
#define GM_D2I(a) ((int32_t)(a))
#define GM_I2D(a) ((double)(a))
#define GM_L2D(a) ((double)(int64_t)(a))
static inline uint32_t gm_fbits(float a) {
    uint32_t b;
    memcpy(&b, &a, 4);
    return b;
}
static inline float gm_bitsf(uint32_t b) {
    float a;
    memcpy(&a, &b, 4);
    return a;
}
#define GM_FBITS(a) gm_fbits(a)
#define GM_BITSF(a) gm_bitsf(a)
#endif

namespace glibc_math {

// ---- sinf / cosf ------------------------------------------------------------------------------------------------------------
template <bool kFma> GM_FN double madd(double a, double b, double c) { return kFma ? GM_DFMA(a, b, c) : GM_DADD(GM_DMUL(a, b), c); }

// sinf_poly (sincosf.h): n even -> sine polynomial of x, n odd -> cosine polynomial; neg = the negated cosine table
template <bool kFma> GM_FN float sincos_poly(double x, double x2, int n, bool neg) {
    if ((n & 1) == 0) {
        const double s1c = -0x1.555545995a603p-3, s2c = 0x1.1107605230bc4p-7, s3c = -0x1.994eb3774cf24p-13;
        const double x3 = GM_DMUL(x, x2);
        const double s1 = madd<kFma>(x2, s3c, s2c);
        const double x7 = GM_DMUL(x3, x2);
        const double s = madd<kFma>(x3, s1c, x);
        return GM_D2F(madd<kFma>(x7, s1, s));
    }
    const double sg = neg ? -1.0 : 1.0; // table[1] holds the same magnitudes with the other sign (exact)
    const double c0 = sg * 0x1p0, c1c = sg * -0x1.ffffffd0c621cp-2, c2c = sg * 0x1.55553e1068f19p-5, c3 = sg * -0x1.6c087e89a359dp-10,
                 c4 = sg * 0x1.99343027bf8c3p-16;
    const double x4 = GM_DMUL(x2, x2);
    const double c2 = madd<kFma>(x2, c4, c3);
    const double c1 = madd<kFma>(x2, c1c, c0);
    const double x6 = GM_DMUL(x4, x2);
    const double c = madd<kFma>(x4, c2c, c1);
    return GM_D2F(madd<kFma>(x6, c2, c));
}
GM_FN double quadrant_sign(int n) { return ((n + 1) & 2) ? -1.0 : 1.0; } // sign[] = {1, -1, -1, 1}
template <bool kFma> GM_FN double reduce_fast(double x, int *np) {
    const double r = GM_DMUL(x, 0x1.45f306dc9c883p+23); // 2/pi * 2^24
    const int n = (GM_D2I(r) + 0x800000) >> 24;
    *np = n;
    return madd<kFma>(-GM_I2D(n), 0x1.921fb54442d18p+0, x); // x - n * pi/2
}
GM_FN uint32_t inv_pio4_word(int i) { // __inv_pio4[i]: 32-bit windows, 8 bits apart, of the bits of 4/pi (0xa2f9836e 4e441529 fc2757d1 f534ddc0 db629599 3c439041)
    const uint64_t w[4] = {0x000000a2f9836e4eull, 0x441529fc2757d1f5ull, 0x34ddc0db6295993cull, 0x4390410000000000ull};
    const int bit = 8 * i; // window = bits [bit, bit + 32) of the 256-bit string w[0..3], counted from its top
    const int q = bit >> 6, r = bit & 63;
    const uint64_t hi = w[q], lo = q + 1 < 4 ? w[q + 1] : 0;
    const uint64_t v = r ? (hi << r) | (lo >> (64 - r)) : hi;
    return (uint32_t)(v >> 32);
}
GM_FN double reduce_large(uint32_t xi, int *np) {
    const int a = (xi >> 26) & 15;
    const int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    res0 = (uint32_t)(xi * inv_pio4_word(a));
    res1 = (uint64_t)xi * inv_pio4_word(a + 4);
    res2 = (uint64_t)xi * inv_pio4_word(a + 8);
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    *np = (int)n;
    return GM_DMUL(GM_L2D((long long)res0), 0x1.921fb54442d18p-62);
}
template <bool kFma> GM_FN float sinf_(float y) {
    const uint32_t xi = GM_FBITS(y), top = (xi >> 20) & 0x7ff;
    double x = GM_F2D(y);
    int n;
    if (top < 0x3f4) {             // |y| < pi/4
        if (top < 0x398) return y; // |y| < 2^-12
        return sincos_poly<kFma>(x, GM_DMUL(x, x), 0, false);
    }
    if (top < 0x42f) { // |y| < 120
        x = reduce_fast<kFma>(x, &n);
        return sincos_poly<kFma>(GM_DMUL(x, quadrant_sign(n)), GM_DMUL(x, x), n, (n & 2) != 0);
    }
    if (top < 0x7f8) {
        const int sign = (int)(xi >> 31);
        x = reduce_large(xi, &n);
        return sincos_poly<kFma>(GM_DMUL(x, quadrant_sign(n + sign)), GM_DMUL(x, x), n, ((n + sign) & 2) != 0);
    }
    return GM_FSUB(y, y); // inf, nan
}
template <bool kFma> GM_FN float cosf_(float y) {
    const uint32_t xi = GM_FBITS(y), top = (xi >> 20) & 0x7ff;
    double x = GM_F2D(y);
    int n;
    if (top < 0x3f4) {
        if (top < 0x398) return 1.0f;
        return sincos_poly<kFma>(x, GM_DMUL(x, x), 1, false);
    }
    if (top < 0x42f) {
        x = reduce_fast<kFma>(x, &n);
        return sincos_poly<kFma>(GM_DMUL(x, quadrant_sign(n + 1)), GM_DMUL(x, x), n ^ 1, ((n + 1) & 2) != 0);
    }
    if (top < 0x7f8) {
        x = reduce_large(xi, &n); // cos is even: the argument's sign plays no part
        return sincos_poly<kFma>(GM_DMUL(x, quadrant_sign(n + 1)), GM_DMUL(x, x), n ^ 1, ((n + 1) & 2) != 0);
    }
    return GM_FSUB(y, y);
}

// ---- atanf / atan2f (fdlibm, binary32) ---------------------------------------------------------------------------------------
GM_FN float atanf_(float x) {
    const uint32_t hx = GM_FBITS(x), ix = hx & 0x7fffffffu;
    const bool neg = (hx >> 31) != 0;
    const float hi3 = GM_BITSF(0x3fc90fdau), lo3 = GM_BITSF(0x33a22168u);
    if (ix >= 0x4c000000u) { // |x| >= 2^25
        if (ix > 0x7f800000u) return GM_FADD(x, x);
        return neg ? GM_FSUB(-hi3, lo3) : GM_FADD(lo3, hi3);
    }
    float hi = 0.f, lo = 0.f;
    int id = -1;
    if (ix < 0x3ee00000u) {             // |x| < 0.4375
        if (ix < 0x31000000u) return x; // |x| < 2^-29
    } else {
        const float ax = GM_BITSF(ix);
        if (ix < 0x3f980000u) {     // |x| < 1.1875
            if (ix < 0x3f300000u) { // 7/16 <= |x| < 11/16
                id = 0, hi = GM_BITSF(0x3eed6338u), lo = GM_BITSF(0x31ac3769u);
                x = GM_FDIV(GM_FSUB(GM_FADD(ax, ax), 1.0f), GM_FADD(ax, 2.0f));
            } else {
                id = 1, hi = GM_BITSF(0x3f490fdau), lo = GM_BITSF(0x33222168u);
                x = GM_FDIV(GM_FSUB(ax, 1.0f), GM_FADD(ax, 1.0f));
            }
        } else if (ix < 0x401c0000u) { // |x| < 2.4375
            id = 2, hi = GM_BITSF(0x3f7b985eu), lo = GM_BITSF(0x33140fb4u);
            x = GM_FDIV(GM_FSUB(ax, 1.5f), GM_FADD(GM_FMUL(ax, 1.5f), 1.0f));
        } else {
            id = 3, hi = hi3, lo = lo3;
            x = GM_FDIV(-1.0f, ax);
        }
    }
    const float z = GM_FMUL(x, x), w = GM_FMUL(z, z);
    float s1 = GM_BITSF(0x3c8569d7u); // aT[10]
    s1 = GM_FADD(GM_FMUL(s1, w), GM_BITSF(0x3d4bda59u));
    s1 = GM_FADD(GM_FMUL(s1, w), GM_BITSF(0x3d886b35u));
    s1 = GM_FADD(GM_FMUL(s1, w), GM_BITSF(0x3dba2e6eu));
    s1 = GM_FADD(GM_FMUL(s1, w), GM_BITSF(0x3e124925u));
    s1 = GM_FADD(GM_FMUL(s1, w), GM_BITSF(0x3eaaaaabu));
    s1 = GM_FMUL(s1, z);
    float s2 = GM_BITSF(0xbd15a221u); // aT[9]
    s2 = GM_FSUB(GM_FMUL(s2, w), GM_BITSF(0x3d6ef16bu));
    s2 = GM_FSUB(GM_FMUL(s2, w), GM_BITSF(0x3d9d8795u));
    s2 = GM_FSUB(GM_FMUL(s2, w), GM_BITSF(0x3de38e38u));
    s2 = GM_FSUB(GM_FMUL(s2, w), GM_BITSF(0x3e4ccccdu));
    s2 = GM_FMUL(s2, w);
    const float xs = GM_FMUL(x, GM_FADD(s1, s2));
    if (id < 0) return GM_FSUB(x, xs);
    const float r = GM_FSUB(hi, GM_FSUB(GM_FSUB(xs, lo), x));
    return neg ? -r : r;
}
GM_FN float atan2f_(float y, float x) {
    const uint32_t hx = GM_FBITS(x), hy = GM_FBITS(y), ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    const float tiny = 1.0e-30f, pi_o_4 = GM_BITSF(0x3f490fdbu), pi_o_2 = GM_BITSF(0x3fc90fdbu), pi = GM_BITSF(0x40490fdbu), pi_lo = GM_BITSF(0xb3bbbd2eu);
    if (ix > 0x7f800000u || iy > 0x7f800000u) return GM_FADD(x, y);
    if (hx == 0x3f800000u) return atanf_(y);
    const int m = (int)((hy >> 31) & 1) | (int)((hx >> 30) & 2);
    if (iy == 0) return m < 2 ? y : (m == 2 ? GM_FADD(pi, tiny) : GM_FSUB(-pi, tiny));
    if (ix == 0) return (hy >> 31) ? GM_FSUB(-pi_o_2, tiny) : GM_FADD(pi_o_2, tiny);
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
            case 0: return GM_FADD(pi_o_4, tiny);
            case 1: return GM_FSUB(-pi_o_4, tiny);
            case 2: return GM_FADD(GM_FMUL(3.0f, pi_o_4), tiny);
            default: return GM_FSUB(GM_FMUL(-3.0f, pi_o_4), tiny);
            }
        }
        switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return GM_FADD(pi, tiny);
        default: return GM_FSUB(-pi, tiny);
        }
    }
    if (iy == 0x7f800000u) return (hy >> 31) ? GM_FSUB(-pi_o_2, tiny) : GM_FADD(pi_o_2, tiny);
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = GM_FADD(pi_o_2, GM_FMUL(0.5f, pi_lo));
    else if ((hx >> 31) && k < -60) z = 0.0f;
    else z = atanf_(GM_BITSF(GM_FBITS(GM_FDIV(y, x)) & 0x7fffffffu));
    switch (m) {
    case 0: return z;
    case 1: return GM_BITSF(GM_FBITS(z) ^ 0x80000000u);
    case 2: return GM_FSUB(pi, GM_FSUB(z, pi_lo));
    default: return GM_FSUB(GM_FSUB(z, pi_lo), pi);
    }
}

} // namespace glibc_math
