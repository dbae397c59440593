// qd_stft.cu -- batched strided-window STFT + magnitude + range threshold (spark_fft, src/fft.rs:12-69).
//
// One team of W/16 threads per window, 256/team windows per CTA, widths 1..4096.  The FFT is our radix-4
// DIT definition (see oracle/quadrs_oracle.c fft_rec; rustfft is not in the reference tree) evaluated
// with non-contracted operations, so it is bit-identical to the oracle.  What differs from the general
// gk_fft kernel is only the schedule:
//   * every thread keeps 16 points in registers and does two radix-4 levels per pass (a radix-2 level
//     and one radix-4 level first when log2 W is odd), so a 4096-point window takes three passes;
//   * the first pass reads the window straight from global memory with coalesced strided loads -- thread
//     t' of a window takes samples t' + (W/16) i' and those are exactly the 16 leaf positions of group
//     digitrev(t') -- decoding raw capture bytes in the same load when the chain has no stage;
//   * passes exchange data through a padded shared-memory image (index + index/16 + index/256) that is
//     conflict-free for the scattered first-pass store and the strided later passes;
//   * the last pass feeds the epilogue from registers: fftshifted bin, glyph index by comparing
//     fl64(re^2 + im^2) with per-glyph thresholds (no square root), magnitudes only when asked for.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "qd_device_math.cuh"
#include "qd_internal.h"
#include "qd_stft_epilogue.cuh"

namespace qd {

constexpr int kStftThreads = 256;

__host__ __device__ constexpr uint32_t pad_idx(uint32_t i) { return i + (i >> 4) + (i >> 8); }
// pad_idx(base + q*i) = pad_idx(base) + pad_off(q, i) for every (base, q, i) the passes below use (a group's base
// has no bits between q and 16q resp. 4q, so nothing carries into the padding terms; scripts/check_stft_pad.py
// enumerates all of them): one padded base per thread and compile-time offsets instead of two shifts per access
// Pitch of one window's padded image (float2): with fewer than 16 threads per window a half warp spans several
// windows, and its 64-bit accesses (thread lt of a window at 17*lt float2 plus a common offset) are free of bank
// conflicts when the pitch is congruent to the team size modulo 16 (W = 64: 68 instead of 69, which put the
// fourth window of a half warp on the banks of the first).  Pitch of one window's glyph row (bytes): W + 16, so
// the byte stores of the 8 windows of a warp (W = 64) fall on different banks and rows stay 16-byte aligned.
__host__ __device__ constexpr uint32_t win_pitch(uint32_t W)
{
    const uint32_t old = W + (W >> 4) + (W >> 8) + 1, tw = W / 16;
    if (tw <= 1 || tw >= 16 || (W & 0x55555555u) == 0) return old; // one window per half warp (or thread), or log2 W odd
    const uint32_t need = pad_idx(W - 1) + 1;
    return need + ((tw + 16 - (need & 15)) & 15);
}
__host__ __device__ constexpr uint32_t glyph_pitch(uint32_t W) { return W + 16; }
__host__ __device__ constexpr uint32_t pad_off(uint32_t q, uint32_t i) { return q * i + ((q * i) >> 4) + ((q * i) >> 8); }

// base-4 digit reversal of x over nd digits
__device__ __forceinline__ uint32_t digitrev4(uint32_t x, int nd)
{
    if (nd <= 0) return 0;
    const uint32_t r = __brev(x) >> (32 - 2 * nd);
    return ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);
}

__device__ __forceinline__ float2 load_elem(const FftArgs &a, uint64_t u, uint32_t n)
{
    if (a.raw) return decode_sample(a.raw, a.raw_fmt, a.raw_first + u * a.in_pitch + n);
    if (a.tail_len && n >= a.W - a.tail_len) return a.tail[u * a.tail_len + (n - (a.W - a.tail_len))];
    return a.in[u * a.in_pitch + n];
}

// N strided elements of one window, with the source format resolved once (MODE -1: cf32 windows)
template <int MODE, int N, typename Place>
__device__ __forceinline__ void load_group_m(const FftArgs &a, uint64_t u, uint32_t first, uint32_t step, float2 *e, Place place)
{
    if (MODE < 0) {
        // the window in the stream, its last tail_len samples in the patch matrix: both bases once per window
        const float2 *__restrict__ win = a.in + u * a.in_pitch;
        const uint32_t n_tail = a.W - a.tail_len; // = W without a patch
        const float2 *__restrict__ tl = a.tail + u * a.tail_len - n_tail;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t n = first + step * i;
            e[place(i)] = *((n >= n_tail ? tl : win) + n);
        }
    } else {
        const uint64_t base = a.raw_first + u * a.in_pitch; // first sample of the window
#pragma unroll
        for (int i = 0; i < N; i++) e[place(i)] = decode_sample_packed<MODE>(a.raw, base + first + step * i, a.one);
    }
}
template <int N, typename Place>
__device__ __forceinline__ void load_group(const FftArgs &a, uint64_t u, uint32_t first, uint32_t step, float2 *e, Place place)
{
    if (!a.raw) return load_group_m<-1, N>(a, u, first, step, e, place);
    switch (a.raw_fmt) {
    case QD_FMT_CF32: return load_group_m<QD_FMT_CF32, N>(a, u, first, step, e, place);
    case QD_FMT_CS8: return load_group_m<QD_FMT_CS8, N>(a, u, first, step, e, place);
    case QD_FMT_CU8: return load_group_m<QD_FMT_CU8, N>(a, u, first, step, e, place);
    default: return load_group_m<QD_FMT_CS16, N>(a, u, first, step, e, place);
    }
}


// two radix-4 levels on 16 points held by one thread: element i sits at position k + q*i of its block;
// level 1 has sub-size q (twiddle index k), level 2 sub-size 4q (twiddle index k + q*c).  tw(j) is the thread's
// j-th twiddle in order of use: j < 3: w(4q, (j+1) k); then w(16q, m (k + q c)) at j = 3 + 3c + (m-1).
template <typename Tw>
__device__ __forceinline__ void levels2(float2 (&e)[16], uint32_t k, uint32_t q, float2 one, Tw tw)
{
    if (k != 0) {
        const float2 w1 = tw(0), w2 = tw(1), w3 = tw(2);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            e[4 * j + 1] = pmul_tw(e[4 * j + 1], w1, one);
            e[4 * j + 2] = pmul_tw(e[4 * j + 2], w2, one);
            e[4 * j + 3] = pmul_tw(e[4 * j + 3], w3, one);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) pradix4(e[4 * j], e[4 * j + 1], e[4 * j + 2], e[4 * j + 3]);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t kp = k + q * c;
        if (kp != 0) {
            e[c + 4] = pmul_tw(e[c + 4], tw(3 + 3 * c), one);
            e[c + 8] = pmul_tw(e[c + 8], tw(4 + 3 * c), one);
            e[c + 12] = pmul_tw(e[c + 12], tw(5 + 3 * c), one);
        }
        pradix4(e[c], e[c + 4], e[c + 8], e[c + 12]);
    }
}
// the j-th twiddle of levels2 out of the natural table T = w(W, .)
__device__ __forceinline__ float2 tw_natural(const float2 *__restrict__ T, uint32_t W, uint32_t k, uint32_t q, int j)
{
    if (j < 3) return __ldg(T + (j + 1) * k * (W / (4 * q)));
    const int c = (j - 3) / 3, m = (j - 3) % 3 + 1;
    return __ldg(T + m * (k + q * c) * (W / (16 * q)));
}

// one radix-4 level on 4 points: element i at position k + q*i, sub-size q
__device__ __forceinline__ void levels1(float2 (&e)[4], uint32_t k, uint32_t q, uint32_t W, const float2 *__restrict__ T, float2 one)
{
    if (k != 0) {
        const uint32_t sc = W / (4 * q);
        e[1] = pmul_tw(e[1], __ldg(T + k * sc), one);
        e[2] = pmul_tw(e[2], __ldg(T + 2 * k * sc), one);
        e[3] = pmul_tw(e[3], __ldg(T + 3 * k * sc), one);
    }
    pradix4(e[0], e[1], e[2], e[3]);
}

// radix-2 innermost level and the first radix-4 level (sub-size 2) on 8 points, log2 W odd
__device__ __forceinline__ void first_odd8(float2 (&e)[8], uint32_t W, const float2 *__restrict__ T, float2 one)
{
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const float2 p = e[2 * a], q = e[2 * a + 1];
        e[2 * a] = padd(p, q);
        e[2 * a + 1] = psub(p, q);
    }
    pradix4(e[0], e[2], e[4], e[6]); // k = 0
    const uint32_t sc = W / 8;      // k = 1: w(8, c)
    e[3] = pmul_tw(e[3], __ldg(T + sc), one);
    e[5] = pmul_tw(e[5], __ldg(T + 2 * sc), one);
    e[7] = pmul_tw(e[7], __ldg(T + 3 * sc), one);
    pradix4(e[1], e[3], e[5], e[7]);
}

// A whole window of N <= 16 bins held by one thread (element i is FFT bin i): glyphs packed into words
template <int N>
__device__ __forceinline__ void emit_window(const FftArgs &a, uint64_t u, const float2 *e)
{
    if (N < 4 || a.mag || !a.use_thr || (reinterpret_cast<uintptr_t>(a.idx) & 3)) {
#pragma unroll
        for (int i = 0; i < N; i++) emit_bin(a, u, N, i, e[i]);
        return;
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(a.idx + static_cast<size_t>(u) * N);
#pragma unroll
    for (int j = 0; j < N / 4; j++) { // word j = display positions 4j .. 4j+3 = bins (4j + N/2 ..) mod N
        const int i = (4 * j + N / 2) & (N - 1);
        o[j] = glyph4_word(a, e[i], e[(i + 1) & (N - 1)], e[(i + 2) & (N - 1)], e[(i + 3) & (N - 1)]);
    }
}

// four bins of a wider window (positions pos, pos + q, ...): staged as bytes in the team's row of shared memory
// (gl), written out 16 bytes per thread once the whole row is there
__device__ __forceinline__ void stage_bins4(const FftArgs &a, uint8_t *gl, uint32_t W, uint32_t pos, uint32_t q, const float2 *e)
{
    uint32_t g[4];
    glyph4(a, e[0], e[1], e[2], e[3], g);
#pragma unroll
    for (int r = 0; r < 4; r++) gl[(pos + q * r + W / 2) & (W - 1)] = static_cast<uint8_t>(g[r]);
}

template <int LOGW, int MINB = 4>
__global__ void __launch_bounds__(kStftThreads, MINB) fk_stft(const __grid_constant__ FftArgs a)
{
    constexpr uint32_t W = 1u << LOGW;
    constexpr int M = LOGW / 2;
    constexpr bool ODD = LOGW & 1;
    constexpr uint32_t TW = W >= 16 ? W / 16 : 1;          // threads per window
    constexpr uint32_t WPC = kStftThreads / TW;            // windows per CTA
    constexpr int FIRST_BITS = ODD ? (LOGW >= 3 ? 3 : 1) : (LOGW >= 4 ? 4 : LOGW);
    constexpr int REST = LOGW - FIRST_BITS;                // log2 of what the later passes still combine
    constexpr int N16 = REST / 4, N4 = (REST % 4) / 2;
    constexpr uint32_t WIN_SMEM = win_pitch(W);
    constexpr uint32_t X_ELEMS = (WPC * WIN_SMEM + 1) & ~1u; // float2 elements, so the glyph rows start 16-byte aligned
    extern __shared__ __align__(16) float2 stft_smem[];

    const uint32_t team = threadIdx.x / TW, lt = threadIdx.x % TW;
    const uint64_t u = static_cast<uint64_t>(blockIdx.x) * WPC + team;
    const bool active = u < a.n_units;
    float2 *x = stft_smem + static_cast<size_t>(team) * WIN_SMEM;
    uint8_t *gl = reinterpret_cast<uint8_t *>(stft_smem + X_ELEMS) + static_cast<size_t>(team) * glyph_pitch(W); // the team's glyph row
    const bool staged = !a.mag && a.use_thr; // index-only output: rows leave through shared memory, 16 bytes per thread
    const float2 *__restrict__ T = a.tw;

    // ---------------- pass 1: global -> registers -> first levels ----------------
    if constexpr (ODD && LOGW >= 3) {
        constexpr uint32_t NG = W / 8;                      // groups of 8 leaf positions
        constexpr uint32_t GPT = NG / TW;                   // groups per thread (2 for W >= 32, 1 for W = 8)
        if (active) {
#pragma unroll
            for (uint32_t z = 0; z < GPT; z++) {
                const uint32_t t2 = lt + z * TW;
                float2 e[8];
                // sample i = r_m + 4*b of the stride-NG comb  ->  leaf offset b + 2*r_m
                load_group<8>(a, u, t2, NG, e, [](int i) { return (i >> 2) | ((i & 3) << 1); });
                first_odd8(e, W, T, a.one);
                if (REST == 0) {
                    emit_window<8>(a, u, e);
                } else {
                    float2 *xg = x + pad_idx(8 * digitrev4(t2, M - 1));
#pragma unroll
                    for (int i = 0; i < 8; i++) xg[i] = e[i];
                }
            }
        }
    } else if constexpr (!ODD && LOGW >= 4) {
        if (active) {
            float2 e[16];
            // sample i = r_{m-1} + 4*r_m of the stride-TW comb  ->  leaf offset r_m + 4*r_{m-1}
            load_group<16>(a, u, lt, TW, e, [](int i) { return (i >> 2) | ((i & 3) << 2); });
            levels2(e, 0, 1, a.one, [&](int j) { return tw_natural(T, W, 0, 1, j); });
            if constexpr (REST == 0) {
                emit_window<16>(a, u, e);
            } else {
                float2 *xg = x + pad_idx(16 * digitrev4(lt, M - 2));
#pragma unroll
                for (int i = 0; i < 16; i++) xg[i] = e[i];
            }
        }
    } else if constexpr (LOGW == 2) {
        if (active) {
            float2 e[4];
#pragma unroll
            for (int i = 0; i < 4; i++) e[i] = load_elem(a, u, i);
            pradix4(e[0], e[1], e[2], e[3]);
            emit_window<4>(a, u, e);
        }
    } else if constexpr (LOGW == 1) {
        if (active) {
            const float2 p = load_elem(a, u, 0), q = load_elem(a, u, 1);
            emit_bin(a, u, W, 0, padd(p, q));
            emit_bin(a, u, W, 1, psub(p, q));
        }
    } else if constexpr (LOGW == 0) {
        if (active) emit_bin(a, u, W, 0, load_elem(a, u, 0));
    }
    if constexpr (REST == 0) return;

    // ---------------- later passes through shared memory ----------------
    uint32_t q = 1u << FIRST_BITS; // points combined so far
#pragma unroll
    for (int pass = 0; pass < N16; pass++) {
        __syncthreads();
        const bool last = (pass == N16 - 1) && N4 == 0;
        if (active) {
            const uint32_t k = lt & (q - 1), blk = lt / q; // one group of 16 per thread
            const uint32_t base = blk * 16 * q + k;
            float2 *xb = x + pad_idx(base);
            float2 e[16];
#pragma unroll
            for (int i = 0; i < 16; i++) e[i] = xb[pad_off(q, i)];
            // the thread's 15 twiddles in order of use, [j][k] (stft_thread_twiddles): neighbouring k are neighbouring
            // words, so a load is 1-2 lines per warp; out of the natural table a warp's 15 loads touched 240 lines
            // in the q = 16 pass of W = 4096 and 96 in the q = 256 pass -- the L1 tag stage was the kernel's limiter
            const float2 *__restrict__ P = a.twp + (q - (1u << FIRST_BITS)); // 15 * (sum of q over the earlier passes)
            if (a.twp) levels2(e, k, q, a.one, [&](int j) { return __ldg(P + j * q + k); });
            else levels2(e, k, q, a.one, [&](int j) { return tw_natural(T, W, k, q, j); });
            if (last) {
                if (staged) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) stage_bins4(a, gl, W, base + q * i, q, e + i);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) emit_bin(a, u, W, base + q * i, e[i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) xb[pad_off(q, i)] = e[i];
            }
        }
        q *= 16;
    }
    if (N4) {
        __syncthreads();
        if (active) {
#pragma unroll
            for (uint32_t z = 0; z < 4; z++) { // four groups of 4 per thread
                const uint32_t gid = lt + z * TW;
                const uint32_t k = gid & (q - 1), blk = gid / q;
                const uint32_t base = blk * 4 * q + k;
                const float2 *xb = x + pad_idx(base);
                float2 e[4];
#pragma unroll
                for (int i = 0; i < 4; i++) e[i] = xb[pad_off(q, i)];
                levels1(e, k, q, W, T, a.one);
                if (staged) {
                    stage_bins4(a, gl, W, base, q, e);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) emit_bin(a, u, W, base + q * i, e[i]);
                }
            }
        }
    }
    if (staged) { // every thread of the team carries 16 consecutive bytes of its window's row out
        __syncthreads();
        if (active) {
            const uint4 row = *reinterpret_cast<const uint4 *>(gl + 16 * lt);
            uint8_t *o = a.idx + static_cast<size_t>(u) * W + 16 * lt;
            if ((reinterpret_cast<uintptr_t>(a.idx) & 15) == 0) {
                *reinterpret_cast<uint4 *>(o) = row;
            } else {
                const uint32_t wd[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
                for (int i = 0; i < 16; i++) o[i] = static_cast<uint8_t>(wd[i / 4] >> (8 * (i & 3)));
            }
        }
    }
}

template <int LOGW, int MINB = 4>
static int launch_stft_k(Chain &c, const FftArgs &fa, uint64_t units)
{
    constexpr uint32_t W = 1u << LOGW;
    constexpr uint32_t TW = W >= 16 ? W / 16 : 1;
    constexpr uint32_t WPC = kStftThreads / TW;
    constexpr int FIRST_BITS = (LOGW & 1) ? (LOGW >= 3 ? 3 : 1) : (LOGW >= 4 ? 4 : LOGW);
    const size_t smem = LOGW == FIRST_BITS ? 0 : static_cast<size_t>((WPC * win_pitch(W) + 1) & ~1u) * sizeof(float2) + static_cast<size_t>(WPC) * glyph_pitch(W);
    if (smem > 48 * 1024)
        QD_CUDA(cudaFuncSetAttribute(fk_stft<LOGW, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const uint64_t grid = (units + WPC - 1) / WPC;
    if (grid > 0x7fffffffull) return set_error(QD_E_INVALID_ARG, "too many windows in one launch");
    fk_stft<LOGW, MINB><<<static_cast<unsigned>(grid), kStftThreads, smem, c.stream>>>(fa);
    QD_LAUNCHED();
    return QD_OK;
}

size_t stft_thread_twiddles(size_t W, const float *tw, float *out)
{
    if (W < 128 || W > 4096 || (W & (W - 1))) return 0;
    int logw = 0;
    while ((size_t(1) << logw) < W) logw++;
    const int first = (logw & 1) ? 3 : 4;
    const int n16 = (logw - first) / 4;
    size_t q = size_t(1) << first, n = 0;
    for (int pass = 0; pass < n16; pass++, q *= 16) {
        for (int j = 0; j < 15; j++)
            for (size_t k = 0; k < q; k++) {
                const size_t c = j < 3 ? 0 : (j - 3) / 3, m = j < 3 ? j + 1 : (j - 3) % 3 + 1;
                const size_t idx = j < 3 ? m * k * (W / (4 * q)) : m * (k + q * c) * (W / (16 * q));
                out[2 * (n + j * q + k)] = tw[2 * idx];
                out[2 * (n + j * q + k) + 1] = tw[2 * idx + 1];
            }
        n += 15 * q;
    }
    return n;
}

// The linear form of the glyph boundaries (FftArgs::use_lin), accepted only when it provably agrees with the
// thresholds outside its band.  With g^(s) = lin_a * sqrt(s) + lin_b in exact arithmetic on the ROUNDED f32
// coefficients, every threshold must lie within eps/2 of its integer: g^(thr[c]) = c + 1 for c < 7, = 8 for the
// panic zone and for max.  The kernel's f32 evaluation is off g^ by at most
//   E = (9 + |b|) * 1.5 * 2^-23  (re^2 + im^2 in f32: 2^-23 relative, halved by the root; sqrt.approx: 2^-23)
//     + 9 * 2^-24                (the rounding of the fused multiply-add; g <= 9 after the clamps)
//     + a * 2^-63                (flushed denormals: |sqrt| error <= 2^-63)
// < eps/4 with eps = (9 + |b|) * 2^-20, so a bin whose g is further than eps from every integer has floor(g) on
// the same side of every threshold as s itself (g^ is increasing).  re^2 + im^2 overflowing to inf in f32 must
// mean "above max": g^(2^127) >= 9.
static void stft_linear_glyphs(FftArgs &fa)
{
    fa.use_lin = 0;
    fa.lin_a = fa.lin_bh = fa.lin_bl = fa.lin_lo = fa.lin_hi = 0.0f;
    if (!fa.use_thr) return;
    const double dist = static_cast<double>(fa.distinction);
    if (!(dist > 0.0) || !std::isfinite(dist)) return;
    const float a = static_cast<float>(1.0 / dist);
    const float b = static_cast<float>(1.0 - static_cast<double>(fa.mn) / dist);
    if (!std::isfinite(a) || !std::isfinite(b) || !(a > 0.0f) || a > 1.0e9f || std::fabs(b) > 8192.0f) return;
    const double A = a, B = b;
    const double eps = (9.0 + std::fabs(B)) * std::ldexp(1.0, -20);
    for (int c = 0; c < 9; c++) {
        const double t = fa.thr[c];
        if (!(t >= 0.0) || !std::isfinite(t)) return; // NaN: the boundary does not exist (max = inf, ...)
        const double target = c < 7 ? c + 1 : 8;
        const double g = A * std::sqrt(t) + B;
        if (t == 0.0 ? g < target - eps / 2 : std::fabs(g - target) > eps / 2) return;
    }
    if (A * std::sqrt(std::ldexp(1.0, 127)) + B < 9.0) return;
    const float lo = B < 0.5 ? static_cast<float>((0.5 - B) / A) : 0.0f;
    const float hi = static_cast<float>((8.5 - B) / A);
    if (!(hi > lo) || !std::isfinite(hi)) return;
    fa.lin_a = a;
    fa.lin_bh = static_cast<float>(B - 0.5 + eps);
    fa.lin_bl = static_cast<float>(B - 0.5 - eps);
    fa.lin_lo = lo;
    fa.lin_hi = hi;
    fa.use_lin = 1;
}

// the glyph boundaries as thresholds on re^2 + im^2 (see FftArgs::thr) and the opaque (1, 1)
void stft_finalize_args(FftArgs &fa)
{
    fa.use_thr = spark_thresholds(fa.mn, fa.mx, fa.thr) ? 1 : 0;
    for (int i = 0; i < 9; i++) {
        uint64_t bits;
        memcpy(&bits, &fa.thr[i], 8);
        fa.thr_hi[i] = static_cast<uint32_t>(bits >> 32);
    }
    fa.one = make_float2(1.0f, 1.0f);
    stft_linear_glyphs(fa);
}

int launch_stft_fast(Chain &c, const FftArgs &fa_in, uint64_t units, bool *handled)
{
    *handled = false;
    if (fa_in.epi != EPI_SPARK || fa_in.window || fa_in.W == 0 || fa_in.W > 4096 || (fa_in.W & (fa_in.W - 1))) return QD_OK;
    FftArgs fa = fa_in;
    fa.n_units = units;
    stft_finalize_args(fa);
    if (!c.glyph_lin) fa.use_lin = 0;
    int logw = 0;
    while ((1u << logw) < fa.W) logw++;
    *handled = true;
    switch (logw) {
    case 0: return launch_stft_k<0>(c, fa, units);
    case 1: return launch_stft_k<1>(c, fa, units);
    case 2: return launch_stft_k<2>(c, fa, units);
    case 3: return launch_stft_k<3>(c, fa, units);
    case 4: return launch_stft_k<4>(c, fa, units);
    case 5: return launch_stft_k<5>(c, fa, units);
    case 6: return launch_stft_k<6>(c, fa, units);
    case 7: return launch_stft_k<7>(c, fa, units);
    case 8: return launch_stft_k<8>(c, fa, units);
    case 9: return launch_stft_k<9>(c, fa, units);
    case 10: return launch_stft_k<10>(c, fa, units);
    case 11: return launch_stft_k<11>(c, fa, units);
    case 12: return c.stft_minb == 3 ? launch_stft_k<12, 3>(c, fa, units) : c.stft_minb == 2 ? launch_stft_k<12, 2>(c, fa, units) : launch_stft_k<12>(c, fa, units);
    }
    *handled = false;
    return QD_OK;
}

} // namespace qd
