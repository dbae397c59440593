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
    // (ax wx - ay wy, ax wy + ay wx), each product and each sum rounded on its own.  The sign of the second
    // product rides on the opaque constant, fl(ay (-wy)) = -fl(ay wy), and the halves of w are swapped by the
    // operand's swizzle: no instruction builds (-wy, wx) (it was a negation and a move per twiddle, both on the
    // FMA pipe that bounds fk_stft)
    const float2 p1 = mul2(make_float2(a.x, a.x), w);                     // (ax wx, ax wy)
    const float2 p2 = mul2(make_float2(a.y, a.y), make_float2(w.y, w.x)); // (ay wy, ay wx)
    return fma2(p2, make_float2(-one.x, one.y), p1);
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

// ---- four glyphs at once through the linear form (FftArgs::use_lin) ----
// fft.rs:53-60 is glyph = clamp(floor((norm - min) / distinction + 1), 0, 8) up to its own f32 roundings, so
// g = lin_a * sqrt(re^2 + im^2) + lin_b in f32 (error bound E, stft_finalize_args) decides every bin whose g is
// further than lin_eps >= 4 E from an integer.  floor(g) without a conversion: (g - 0.5 + eps) and (g - 0.5 - eps)
// are both added to 1.5 * 2^23, which rounds them to integers in the low mantissa bits; equal results = same
// integer on both sides of the band = decided.  A NaN propagates through the clamps and compares unequal.
// 8 instructions per bin (FMUL, FFMA, MUFU.SQRT, 2 FMNMX, FFMA2, FADD2, FSETP) against ~40 for the thresholds.
__device__ __forceinline__ float min_nan(float a, float b)
{
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float max_nan(float a, float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// low byte of the result = glyph; undecided |= the bin needs the thresholds
__device__ __forceinline__ uint32_t glyph_lin(const FftArgs &a, float2 v, bool &undecided)
{
    const float s = fmaf(v.y, v.y, v.x * v.x);
    const float r = min_nan(max_nan(sqrt_approx(s), a.lin_lo), a.lin_hi);
    const float2 hl = fma2(make_float2(r, r), make_float2(a.lin_a, a.lin_a), make_float2(a.lin_bh, a.lin_bl));
    const float2 m = add2(hl, make_float2(12582912.0f, 12582912.0f));
    undecided |= !(m.x == m.y);
    return __float_as_uint(m.x);
}
// the rare group with an undecided bin (or no linear form): the thresholds, out of line; byte i = glyph of v_i
static __device__ __noinline__ uint32_t glyph4_by_threshold(const FftArgs &a, float2 v0, float2 v1, float2 v2, float2 v3)
{
    return static_cast<uint32_t>(glyph_fast(a, v0)) | static_cast<uint32_t>(glyph_fast(a, v1)) << 8 |
           static_cast<uint32_t>(glyph_fast(a, v2)) << 16 | static_cast<uint32_t>(glyph_fast(a, v3)) << 24;
}
// glyphs of four bins: the low byte of g[i]
__device__ __forceinline__ void glyph4(const FftArgs &a, float2 v0, float2 v1, float2 v2, float2 v3, uint32_t (&g)[4])
{
    bool undecided = !a.use_lin;
    g[0] = glyph_lin(a, v0, undecided);
    g[1] = glyph_lin(a, v1, undecided);
    g[2] = glyph_lin(a, v2, undecided);
    g[3] = glyph_lin(a, v3, undecided);
    if (undecided) {
        const uint32_t w = glyph4_by_threshold(a, v0, v1, v2, v3);
        g[0] = w;
        g[1] = w >> 8;
        g[2] = w >> 16;
        g[3] = w >> 24;
    }
}
// ... packed into one word, byte i = glyph of v_i
__device__ __forceinline__ uint32_t glyph4_word(const FftArgs &a, float2 v0, float2 v1, float2 v2, float2 v3)
{
    uint32_t g[4];
    glyph4(a, v0, v1, v2, v3, g);
    return __byte_perm(__byte_perm(g[0], g[1], 0x0040), __byte_perm(g[2], g[3], 0x0040), 0x5410);
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
