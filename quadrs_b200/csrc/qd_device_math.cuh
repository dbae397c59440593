// qd_device_math.cuh -- device-side arithmetic that must reproduce the reference bit for bit.
//
// Every helper here uses the *_rn intrinsics, which nvcc never contracts into FMA, so the rounding
// sequence is exactly the reference's (Rust never contracts either).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace qd {

// ---- FileFormat::to_f32, lib.rs:241-255 ------------------------------------------------------
// The quotient must be the correctly rounded IEEE quotient; multiplying by a reciprocal alone is not
// bit-exact.  div_exact gets it without the IEEE divide sequence: q0 = x*c, r = fma(-q0, den, x),
// q = fma(r, c, q0) with c = fl(1/den) -- verified exhaustively on the host for every i8 / 127, u8 / 255
// and i16 / 65535 (tests/test_decode_trick.py).
__device__ __forceinline__ float div_exact(float x, float den, float c)
{
    const float q0 = __fmul_rn(x, c);
    const float r = __fmaf_rn(-q0, den, x);
    return __fmaf_rn(r, c, q0);
}
__device__ __forceinline__ float dec_s8(int v) { return div_exact(static_cast<float>(v), 127.0f, 1.0f / 127.0f); }
__device__ __forceinline__ float dec_u8(unsigned v) { return __fsub_rn(div_exact(static_cast<float>(v), 255.0f, 1.0f / 255.0f), 127.5f); }
__device__ __forceinline__ float dec_s16(int v) { return __fsub_rn(div_exact(static_cast<float>(v), 65535.0f, 1.0f / 65535.0f), 32767.5f); }

// One sample (I first, Q second: lib.rs:234-237) at index i of a raw byte stream.
__device__ __forceinline__ float2 decode_sample(const uint8_t *__restrict__ raw, int fmt, uint64_t i)
{
    switch (fmt) {
    case 0: { // cf32: little-endian bit copy (NaN payloads preserved)
        const float2 v = __ldg(reinterpret_cast<const float2 *>(raw) + i);
        return v;
    }
    case 1: { // cs8
        const char2 v = __ldg(reinterpret_cast<const char2 *>(raw) + i);
        return make_float2(dec_s8(v.x), dec_s8(v.y));
    }
    case 2: { // cu8
        const uchar2 v = __ldg(reinterpret_cast<const uchar2 *>(raw) + i);
        return make_float2(dec_u8(v.x), dec_u8(v.y));
    }
    default: { // cs16
        const short2 v = __ldg(reinterpret_cast<const short2 *>(raw) + i);
        return make_float2(dec_s16(v.x), dec_s16(v.y));
    }
    }
}

// The same values without an integer->float conversion (I2F runs on the 16-lane conversion pipe): a byte or half
// word dropped into the mantissa of 2^23 by PRMT is the float 2^23 + v, one exact subtraction leaves v, and the
// quotient / offset steps are the packed forms of div_exact (each half individually rounded: identical bits).
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b)
{
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
// (one = (1, 1) passed in by the caller as a kernel parameter: opaque to ptxas, so x * 1 + c is not folded)
template <int FMT>
__device__ __forceinline__ float2 decode_sample_packed(const uint8_t *__restrict__ raw, uint64_t i, float2 one)
{
    if (FMT == 0) return __ldg(reinterpret_cast<const float2 *>(raw) + i); // cf32: bit copy
    float2 n;
    float den, c, off;
    if (FMT == 3) { // cs16
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(raw) + i) ^ 0x80008000u;
        n = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410)) - 8421376.0f,
                        __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432)) - 8421376.0f); // -(2^23 + 32768): exact
        den = 65535.0f, c = 1.0f / 65535.0f, off = -32767.5f;
    } else {
        const uint32_t h = static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short *>(raw) + i)) ^ (FMT == 1 ? 0x8080u : 0u);
        const float k = FMT == 1 ? 8388736.0f : 8388608.0f; // 2^23 + 128 (cs8, bytes flipped to offset binary) or 2^23
        n = make_float2(__uint_as_float(__byte_perm(h, 0x4B000000u, 0x7440)) - k, __uint_as_float(__byte_perm(h, 0x4B000000u, 0x7441)) - k);
        den = FMT == 1 ? 127.0f : 255.0f, c = FMT == 1 ? 1.0f / 127.0f : 1.0f / 255.0f, off = -127.5f;
    }
    const float2 c2 = make_float2(c, c);
    // cu8 / cs16 (lib.rs:252-253): fl(fl(x / den) - off) equals ONE fused fl(x * fl(1/den) - off) for every u8 / 255
    // and every i16 / 65535 -- the result's ulp (2^-17, 2^-9) is so much coarser than the quotient's that neither
    // the reciprocal's error nor the first rounding ever moves it across a rounding boundary; checked exhaustively
    // in exact rational arithmetic by tests/test_decode_trick.py.  One packed FMA instead of four packed operations.
    if (FMT != 1) return f2_fma(n, c2, make_float2(off, off));
    const float2 q0 = f2_mul(n, c2);
    const float2 r = f2_fma(q0, make_float2(-den, -den), n);
    return f2_fma(r, c2, q0); // cs8, lib.rs:251: the correctly rounded quotient (div_exact)
}

// ---- cos/sin of an f64 phase, rounded to f32 (shift.rs:50, gen.rs:41) ---------------------------
// theta is the reference's own f64 phase.  Reduction by 2 pi/256 is exact (the first FMA's result is
// representable because |r| < 2^-5 while 2 pi/256 has its last bit at 2^-58); the table is
// double-double, so the f64 result carries < 0.51 ulp before the final rounding to f32 -- the same
// class of error as glibc's sin/cos, hence the same f32 except on ~2^-28 of inputs.
__device__ __forceinline__ void sincos_f64(double theta, const double *__restrict__ tab, double &cd, double &sd)
{
    const double K = 40.743665431525205956834243423364;   // 256 / (2 pi)
    const double C1 = 6.283185307179586 / 256.0;           // fl64(2 pi) / 256, exact scaling
    const double C2 = 2.4492935982947064e-16 / 256.0;      // (2 pi - fl64(2 pi)) / 256
    const double kd = rint(__dmul_rn(theta, K));
    double r = fma(-kd, C1, theta);
    r = fma(-kd, C2, r);
    const long long ki = static_cast<long long>(kd);
    const double2 *tp = reinterpret_cast<const double2 *>(tab) + 2 * (ki & 127); // first half turn, negated for the second
    const double2 tcos = __ldg(tp), tsin = __ldg(tp + 1);
    const double sg = (ki & 128) ? -1.0 : 1.0;
    const double4 t = make_double4(sg * tcos.x, sg * tcos.y, sg * tsin.x, sg * tsin.y); // {ch, cl, sh, sl}
    const double r2 = __dmul_rn(r, r);
    double ps = fma(r2, -1.0 / 5040.0, 1.0 / 120.0);
    ps = fma(r2, ps, -1.0 / 6.0);
    ps = fma(__dmul_rn(r, r2), ps, r); // sin r
    double pc = fma(r2, -1.0 / 720.0, 1.0 / 24.0);
    pc = fma(r2, pc, -0.5);
    pc = __dmul_rn(r2, pc); // cos r - 1
    double tc = fma(t.x, pc, t.y);
    tc = fma(-t.z, ps, tc);
    cd = __dadd_rn(t.x, tc);
    double ts = fma(t.z, pc, t.w);
    ts = fma(t.x, ps, ts);
    sd = __dadd_rn(t.z, ts);
}

// Same arithmetic with the nine f64 constants read from a kernel parameter (constant bank operands of
// DFMA/DMUL) instead of being rebuilt in uniform registers at every use.
struct SinCosK {
    double K, C1, C2, s7, s5, s3, c6, c4, c2;
};
inline SinCosK make_sincos_k()
{
    SinCosK k;
    k.K = 40.743665431525205956834243423364;
    k.C1 = 6.283185307179586 / 256.0;
    k.C2 = 2.4492935982947064e-16 / 256.0;
    k.s7 = -1.0 / 5040.0;
    k.s5 = 1.0 / 120.0;
    k.s3 = -1.0 / 6.0;
    k.c6 = -1.0 / 720.0;
    k.c4 = 1.0 / 24.0;
    k.c2 = -0.5;
    return k;
}
// HALF: read only the first half turn of the table and flip signs for the second (see below); the kernels whose
// tiles are staged in shared memory gain 10 - 15 % from it, the long-filter kernels (tiles read through L1) lose 4 %
template <bool HALF = true>
__device__ __forceinline__ void sincos_f64k(double theta, const double *__restrict__ tab, const SinCosK &k, double &cd,
                                            double &sd)
{
    // kd = rint(theta * K) and its low bits from one addition: 1.5 * 2^52 + p has ulp 1, so the add rounds p to the
    // nearest-even integer (= rint; |p| < 2^43 here) and leaves that integer in the low mantissa bits
    const double t = __dadd_rn(__dmul_rn(theta, k.K), 6755399441055744.0);
    const double kd = __dadd_rn(t, -6755399441055744.0);
    double r = fma(-kd, k.C1, theta);
    r = fma(-kd, k.C2, r);
    // Only the first half turn of the table is read: e^{i(x + pi)} = -e^{ix}, and the second half of the table IS the
    // negated first half (sincos_table).  The gather is the exact mixer's bottleneck -- every lane of a warp reads
    // its own entry, and L1 looks up one 128-byte line at a time -- so halving the lines the lanes can spread over
    // is worth more than the four sign flips cost: config 2 EXACT 257 -> 324 Gsamples/s with 128 entries (327 with 64).
    const int kf = __double2loint(t);
    const int ki = kf & (HALF ? 127 : 255);
    // one 256-bit load per entry {ch, cl, sh, sl} (LDG.E.256, sm_100): half the L1 lookups of two 128-bit loads
    double2 tcos, tsin;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
        : "=d"(tcos.x), "=d"(tcos.y), "=d"(tsin.x), "=d"(tsin.y)
        : "l"(tab + 4 * ki));
    if (HALF) {
        const int neg = (kf & 128) << 24; // sign-bit mask
        tcos.x = __hiloint2double(__double2hiint(tcos.x) ^ neg, __double2loint(tcos.x));
        tcos.y = __hiloint2double(__double2hiint(tcos.y) ^ neg, __double2loint(tcos.y));
        tsin.x = __hiloint2double(__double2hiint(tsin.x) ^ neg, __double2loint(tsin.x));
        tsin.y = __hiloint2double(__double2hiint(tsin.y) ^ neg, __double2loint(tsin.y));
    }
    const double r2 = __dmul_rn(r, r);
    double ps = fma(r2, k.s7, k.s5);
    ps = fma(r2, ps, k.s3);
    ps = fma(__dmul_rn(r, r2), ps, r); // sin r
    double pc = fma(r2, k.c6, k.c4);
    pc = fma(r2, pc, k.c2);
    pc = __dmul_rn(r2, pc); // cos r - 1
    double tc = fma(tcos.x, pc, tcos.y);
    tc = fma(-tsin.x, ps, tc);
    cd = __dadd_rn(tcos.x, tc);
    double ts = fma(tsin.x, pc, tsin.y);
    ts = fma(tcos.x, ps, ts);
    sd = __dadd_rn(tsin.x, ts);
}

__device__ __forceinline__ float2 phasor_exact(uint64_t n, double ratio, const double *__restrict__ tab)
{
    // shift.rs:49: (off + i) as f64 * self.ratio
    const double place = __dmul_rn(__ull2double_rn(n), ratio);
    double c, s;
    sincos_f64(place, tab, c, s);
    return make_float2(static_cast<float>(c), static_cast<float>(s));
}

// ---- streaming global loads that do not allocate in L1: tiles read straight from global memory pass through once,
// and must not evict the sin/cos table (8 KB, read twice per sample by the exact mixer) from the little L1 that is
// left beside the kernel's shared memory
__device__ __forceinline__ uint4 ldg_stream_v4(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_stream_v2(const void *p)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// ---- PTX helpers: shared-window addresses, mbarriers, 1-D bulk async copies (TMA)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
// 1-D bulk async copy global -> shared (TMA); completion is signalled on the mbarrier as bytes land
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// bulk prefetch of a byte range into L2 (no registers, no shared memory)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ---- packed f32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2), each half an individually rounded IEEE op
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

// buf[i] *= mul  (num-complex 0.4.6 MulAssign): re' = re*c - im*s ; im' = im*c + re*s
__device__ __forceinline__ float2 cmul_exact(float2 a, float2 m)
{
    return make_float2(__fsub_rn(__fmul_rn(a.x, m.x), __fmul_rn(a.y, m.y)),
                       __fadd_rn(__fmul_rn(a.y, m.x), __fmul_rn(a.x, m.y)));
}

// Complex * Complex as num-complex Mul: (ar*br - ai*bi, ar*bi + ai*br) -- used by the FFT twiddles.
__device__ __forceinline__ float2 cmul_tw(float2 a, float2 b)
{
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }

// One radix-4 DIT butterfly of our FFT definition (see oracle/quadrs_oracle.c fft_rec):
// inputs already twiddled; outputs X[k], X[k+q], X[k+2q], X[k+3q].
__device__ __forceinline__ void radix4(float2 &t0, float2 &t1, float2 &t2, float2 &t3)
{
    const float2 s0 = cadd(t0, t2), s1 = csub(t0, t2), s2 = cadd(t1, t3), s3 = csub(t1, t3);
    t0 = cadd(s0, s2);
    t1 = make_float2(__fadd_rn(s1.x, s3.y), __fsub_rn(s1.y, s3.x)); // s1 - i*s3
    t2 = csub(s0, s2);
    t3 = make_float2(__fsub_rn(s1.x, s3.y), __fadd_rn(s1.y, s3.x)); // s1 + i*s3
}

// Complex::norm = re.hypot(im) -> glibc hypotf = (float)sqrt((double)x*x + (double)y*y) with the
// inf/nan preamble of sysdeps/ieee754/flt-32/e_hypotf.c.  The f64 products are exact.
__device__ __forceinline__ float hypot_exact(float x, float y)
{
    if (!isfinite(x) || !isfinite(y)) {
        if (isinf(x) || isinf(y)) return __int_as_float(0x7f800000);
        return x + y;
    }
    const double a = x, b = y;
    return static_cast<float>(__dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b))));
}

// fft.rs:53-60.  Returns 0..8, or 9 where the reference would index graph[7] and panic.
__device__ __forceinline__ int glyph_index(float norm, float mn, float mx, float distinction)
{
    if (norm < mn) return 0;
    if (norm >= mx) return 8;
    const float q = __fdiv_rn(__fsub_rn(norm, mn), distinction);
    if (!(q > 0.0f)) return 1; // `as usize` saturates NaN / negatives to 0
    if (q >= 7.0f) return 9;
    return 1 + static_cast<int>(q);
}

} // namespace qd
