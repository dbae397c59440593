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
#include <cstring>

#include "qd_device_math.cuh"
#include "qd_internal.h"
#include "qd_stft_epilogue.cuh"

namespace qd {

constexpr int kStftThreads = 256;

__host__ __device__ constexpr uint32_t pad_idx(uint32_t i) { return i + (i >> 4) + (i >> 8); }

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
// level 1 has sub-size q (twiddle index k), level 2 sub-size 4q (twiddle index k + q*c)
__device__ __forceinline__ void levels2(float2 (&e)[16], uint32_t k, uint32_t q, uint32_t W, const float2 *__restrict__ T, float2 one)
{
    if (k != 0) {
        const uint32_t sc = W / (4 * q);
        const float2 w1 = __ldg(T + k * sc), w2 = __ldg(T + 2 * k * sc), w3 = __ldg(T + 3 * k * sc);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            e[4 * j + 1] = pmul_tw(e[4 * j + 1], w1, one);
            e[4 * j + 2] = pmul_tw(e[4 * j + 2], w2, one);
            e[4 * j + 3] = pmul_tw(e[4 * j + 3], w3, one);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) pradix4(e[4 * j], e[4 * j + 1], e[4 * j + 2], e[4 * j + 3]);
    const uint32_t sc2 = W / (16 * q);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t kp = k + q * c;
        if (kp != 0) {
            e[c + 4] = pmul_tw(e[c + 4], __ldg(T + kp * sc2), one);
            e[c + 8] = pmul_tw(e[c + 8], __ldg(T + 2 * kp * sc2), one);
            e[c + 12] = pmul_tw(e[c + 12], __ldg(T + 3 * kp * sc2), one);
        }
        pradix4(e[c], e[c + 4], e[c + 8], e[c + 12]);
    }
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
    uint32_t w[N >= 4 ? N / 4 : 1];
#pragma unroll
    for (int j = 0; j < N / 4; j++) w[j] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const int b = (i + N / 2) & (N - 1);
        w[b / 4] |= static_cast<uint32_t>(glyph_fast(a, e[i])) << (8 * (b & 3));
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(a.idx + static_cast<size_t>(u) * N);
#pragma unroll
    for (int j = 0; j < N / 4; j++) o[j] = w[j];
}

// 16 bins of a wider window: staged as bytes in the team's row of shared memory (gl), written out by
// store_row once the whole row is there
__device__ __forceinline__ void stage_bin(const FftArgs &a, uint8_t *gl, uint32_t W, uint32_t pos, float2 v)
{
    gl[(pos + W / 2) & (W - 1)] = static_cast<uint8_t>(glyph_fast(a, v));
}

template <int LOGW>
__global__ void __launch_bounds__(kStftThreads, 4) fk_stft(const __grid_constant__ FftArgs a)
{
    constexpr uint32_t W = 1u << LOGW;
    constexpr int M = LOGW / 2;
    constexpr bool ODD = LOGW & 1;
    constexpr uint32_t TW = W >= 16 ? W / 16 : 1;          // threads per window
    constexpr uint32_t WPC = kStftThreads / TW;            // windows per CTA
    constexpr int FIRST_BITS = ODD ? (LOGW >= 3 ? 3 : 1) : (LOGW >= 4 ? 4 : LOGW);
    constexpr int REST = LOGW - FIRST_BITS;                // log2 of what the later passes still combine
    constexpr int N16 = REST / 4, N4 = (REST % 4) / 2;
    constexpr uint32_t WIN_SMEM = pad_idx(W) + 1;
    constexpr uint32_t X_ELEMS = (WPC * WIN_SMEM + 1) & ~1u; // float2 elements, so the glyph rows start 16-byte aligned
    extern __shared__ __align__(16) float2 stft_smem[];

    const uint32_t team = threadIdx.x / TW, lt = threadIdx.x % TW;
    const uint64_t u = static_cast<uint64_t>(blockIdx.x) * WPC + team;
    const bool active = u < a.n_units;
    float2 *x = stft_smem + static_cast<size_t>(team) * WIN_SMEM;
    uint8_t *gl = reinterpret_cast<uint8_t *>(stft_smem + X_ELEMS) + static_cast<size_t>(team) * W; // the team's glyph row
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
                    const uint32_t g = digitrev4(t2, M - 1);
#pragma unroll
                    for (int i = 0; i < 8; i++) x[pad_idx(8 * g + i)] = e[i];
                }
            }
        }
    } else if constexpr (!ODD && LOGW >= 4) {
        if (active) {
            float2 e[16];
            // sample i = r_{m-1} + 4*r_m of the stride-TW comb  ->  leaf offset r_m + 4*r_{m-1}
            load_group<16>(a, u, lt, TW, e, [](int i) { return (i >> 2) | ((i & 3) << 2); });
            levels2(e, 0, 1, W, T, a.one);
            if constexpr (REST == 0) {
                emit_window<16>(a, u, e);
            } else {
                const uint32_t g = digitrev4(lt, M - 2);
#pragma unroll
                for (int i = 0; i < 16; i++) x[pad_idx(16 * g + i)] = e[i];
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
            float2 e[16];
#pragma unroll
            for (int i = 0; i < 16; i++) e[i] = x[pad_idx(base + q * i)];
            levels2(e, k, q, W, T, a.one);
            if (last) {
                if (staged) {
#pragma unroll
                    for (int i = 0; i < 16; i++) stage_bin(a, gl, W, base + q * i, e[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) emit_bin(a, u, W, base + q * i, e[i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) x[pad_idx(base + q * i)] = e[i];
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
                float2 e[4];
#pragma unroll
                for (int i = 0; i < 4; i++) e[i] = x[pad_idx(base + q * i)];
                levels1(e, k, q, W, T, a.one);
                if (staged) {
#pragma unroll
                    for (int i = 0; i < 4; i++) stage_bin(a, gl, W, base + q * i, e[i]);
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

template <int LOGW>
static int launch_stft_k(Chain &c, const FftArgs &fa, uint64_t units)
{
    constexpr uint32_t W = 1u << LOGW;
    constexpr uint32_t TW = W >= 16 ? W / 16 : 1;
    constexpr uint32_t WPC = kStftThreads / TW;
    constexpr int FIRST_BITS = (LOGW & 1) ? (LOGW >= 3 ? 3 : 1) : (LOGW >= 4 ? 4 : LOGW);
    const size_t smem = LOGW == FIRST_BITS ? 0 : static_cast<size_t>((WPC * (pad_idx(W) + 1) + 1) & ~1u) * sizeof(float2) + static_cast<size_t>(WPC) * W;
    if (smem > 48 * 1024)
        QD_CUDA(cudaFuncSetAttribute(fk_stft<LOGW>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const uint64_t grid = (units + WPC - 1) / WPC;
    if (grid > 0x7fffffffull) return set_error(QD_E_INVALID_ARG, "too many windows in one launch");
    fk_stft<LOGW><<<static_cast<unsigned>(grid), kStftThreads, smem, c.stream>>>(fa);
    QD_LAUNCHED();
    return QD_OK;
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
}

int launch_stft_fast(Chain &c, const FftArgs &fa_in, uint64_t units, bool *handled)
{
    *handled = false;
    if (fa_in.epi != EPI_SPARK || fa_in.window || fa_in.W == 0 || fa_in.W > 4096 || (fa_in.W & (fa_in.W - 1))) return QD_OK;
    FftArgs fa = fa_in;
    fa.n_units = units;
    stft_finalize_args(fa);
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
    case 12: return launch_stft_k<12>(c, fa, units);
    }
    *handled = false;
    return QD_OK;
}

} // namespace qd
