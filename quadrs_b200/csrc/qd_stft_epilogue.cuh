// qd_stft_epilogue.cuh -- the pieces of the STFT that more than one kernel uses: the leaf order of our radix-4 FFT
// definition and the magnitude / glyph epilogue of spark_fft (src/fft.rs:48-60).
#pragma once

#include "qd_device_math.cuh"
#include "qd_internal.h"

namespace qd {

// ---- FFT of one unit in shared memory: our radix-4 DIT definition (oracle/quadrs_oracle.c fft_rec) ----
// Leaf position of natural index n: radix-4 digits of n, least significant first, become the most
// significant digits of the position; a leftover top bit (odd log2 W) is the position's bit 0.
__device__ __forceinline__ uint32_t leaf_position(uint32_t n, uint32_t W, int n_r4, bool odd)
{
    uint32_t p = 0, span = W;
    for (int d = 0; d < n_r4; d++) {
        span >>= 2;
        p += (n & 3u) * span;
        n >>= 2;
    }
    if (odd) p += n & 1u;
    return p;
}

// ---- the FFT's complex arithmetic on packed (re, im) pairs.  Every half is the same individually rounded
// operation the scalar form in qd_device_math.cuh performs (a - b = a + (-b) exactly; the product pairs are
// summed through an FMA by an opaque 1.0 so that ptxas cannot contract a multiply into the add).
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return add2(a, b); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { return fma2(b, make_float2(-1.0f, -1.0f), a); }
__device__ __forceinline__ float2 pmul_tw(float2 a, float2 w, float2 one)
{
    const float2 p1 = mul2(make_float2(a.x, a.x), w);                      // (ax wx, ax wy)
    const float2 p2 = mul2(make_float2(a.y, a.y), make_float2(-w.y, w.x)); // (-ay wy, ay wx)
    return fma2(p2, one, p1);
}
__device__ __forceinline__ void pradix4(float2 &t0, float2 &t1, float2 &t2, float2 &t3)
{
    const float2 s0 = padd(t0, t2), s1 = psub(t0, t2), s2 = padd(t1, t3), s3 = psub(t1, t3);
    t0 = padd(s0, s2);
    t1 = padd(s1, make_float2(s3.y, -s3.x)); // s1 - i*s3
    t2 = psub(s0, s2);
    t3 = padd(s1, make_float2(-s3.y, s3.x)); // s1 + i*s3
}

// glyph index of one bin from the thresholds on s = fl64(re^2 + im^2) (spark_thresholds), every case
static __device__ __noinline__ int glyph_by_threshold(const FftArgs &a, float2 v)
{
    const double x = v.x, y = v.y;
    double s = fma(x, x, __dmul_rn(y, y)); // = fl64(x^2 + y^2): both squares are exact in f64
    if (s != s) { // NaN in, or inf - inf: hypotf gives inf if either part is infinite, else NaN
        if (isinf(v.x) || isinf(v.y)) s = __longlong_as_double(0x7ff0000000000000ll);
        else return glyph_index(__int_as_float(0x7fc00000), a.mn, a.mx, a.distinction);
    }
    // r = #{c < 7 : s >= thr[c]} by bisection (thr is non-decreasing; NaN entries never compare true)
    const bool h3 = s >= a.thr[3];
    const bool h1 = s >= (h3 ? a.thr[5] : a.thr[1]);
    const double t0 = h3 ? (h1 ? a.thr[6] : a.thr[4]) : (h1 ? a.thr[2] : a.thr[0]);
    const int r = (h3 ? 4 : 0) + (h1 ? 2 : 0) + (s >= t0 ? 1 : 0);
    return s >= a.thr[8] ? 8 : (s >= a.thr[7] ? 9 : r);
}

// The same decisions on the HIGH WORD of s alone: for finite s >= 0 and a threshold t >= 0 (or NaN),
// s >= t  <=>  hi(s) >= hi(t) unless the two high words are equal.  Any compared threshold whose high word
// equals hi(s), and every non-finite s, goes through the full comparison above.
__device__ __forceinline__ int glyph_fast(const FftArgs &a, float2 v)
{
    const double x = v.x, y = v.y;
    const double s = fma(x, x, __dmul_rn(y, y));
    const uint32_t sh = static_cast<uint32_t>(__double2hiint(s));
    // (all nine high words once, then selects between registers: an index that depends on a comparison would be a
    // constant-bank load per use)
    const uint32_t t0 = a.thr_hi[0], t1 = a.thr_hi[1], t2 = a.thr_hi[2], t3 = a.thr_hi[3], t4 = a.thr_hi[4], t5 = a.thr_hi[5],
                   t6 = a.thr_hi[6];
    const uint32_t c3 = t3;
    const bool h3 = sh >= c3;
    const uint32_t c1 = h3 ? t5 : t1;
    const bool h1 = sh >= c1;
    const uint32_t c0 = h3 ? (h1 ? t6 : t4) : (h1 ? t2 : t0);
    const int r = (h3 ? 4 : 0) + (h1 ? 2 : 0) + (sh >= c0 ? 1 : 0);
    const uint32_t c8 = a.thr_hi[8], c7 = a.thr_hi[7];
    int g = sh >= c8 ? 8 : (sh >= c7 ? 9 : r);
    if (sh == c3 || sh == c1 || sh == c0 || sh == c8 || sh == c7 || sh >= 0x7ff00000u) g = glyph_by_threshold(a, v);
    if (g == 9) *a.panic_flag = 1;
    return g;
}

// glyph index (and optional magnitude) of one output bin, fft.rs:48-60: the general form
__device__ __forceinline__ void emit_bin(const FftArgs &a, uint64_t u, uint32_t W, uint32_t pos, float2 v)
{
    const uint32_t b = W > 1 ? ((pos + W / 2) & (W - 1)) : 0; // display order: bins W/2..W-1 then 0..W/2-1
    const size_t o = static_cast<size_t>(u) * W + b;
    int g;
    if (a.mag || !a.use_thr) {
        const float norm = hypot_exact(v.x, v.y);
        if (a.mag) a.mag[o] = norm;
        g = glyph_index(norm, a.mn, a.mx, a.distinction);
        if (g == 9) *a.panic_flag = 1;
    } else {
        g = glyph_fast(a, v);
    }
    a.idx[o] = static_cast<uint8_t>(g);
}


} // namespace qd
