// qd_tcfir.cu -- fk_tcfir: decode + NCO mix + decimating FIR on the 5th-generation tensor cores (tcgen05.mma,
// accumulators in TMEM), FAST arithmetic, cs8 captures.
//
// What it replaces: LowPass::read_at + convolve behind Shift (filter.rs:54-124, shift.rs:46-54) for the outputs whose
// filter is not truncated, i.e. y[g] = sum_j f[j] * x[n] * e^{i ratio n}, n = g*D + i0 + j.  The truncated outputs at
// the end of every read are patched afterwards by fk_tail (qd_fast.cu), which always uses the exact arithmetic.
//
// Formulation.  The capture is cut into ROWS of 64 samples at ABSOLUTE positions (row b = samples [64b, 64b + 64)), so
// nothing below depends on where a launch, host segment or shard starts.  Writing n = 64b + kk,
//     y[g] = sum_b e^{i ratio 64 b} * P[b][g - OPR*b - c0],      P[b][i'] = sum_kk x[64b + kk] * G[kk][i']
// with OPR = 64/D outputs starting per row and G[kk][i'] = f[kk - (c0 + i') D - i0] / 127 * e^{i ratio kk} (zero outside
// the filter): the frequency translation is folded into complex taps, the raw int8 I/Q pairs of a row ARE the A
// operand (K = 128 reals, exact in f16), and a row touches NOUT = OPR + (L-1)/D outputs.  One tcgen05.mma tile is
// 128 rows x N columns: the columns hold (re, im) of the NOUT partial outputs twice -- taps split as hi + lo*2^-11 in
// f16 so the product carries 22 tap bits -- and the epilogue adds the two halves, rotates a row's partials by the
// row phasor (one f64 sin/cos per row, from the table the exact mixer uses) and sums the 1 + (NOUT-1)/OPR rows that
// meet in an output, in ascending row order.
//
// Roles (one persistent CTA per SM, 19 warps): warps 0-7 epilogue (two groups of four taking alternate tiles, TMEM
// lane quarter = warp & 3); warp 8 TMEM allocation + the single MMA-issuing thread; warp 9 one thread issuing 16 KB
// bulk copies (TMA) of the tiles' raw bytes into a ring of up to 8 slots (~100 KB of loads in flight per SM); warp 10
// the tiles' f64 phase anchors; warps 11-18 converters: raw bytes from the ring, int8 -> f16 by PRMT into the mantissa of 1024 (no
// I2F), 128B-swizzled K-major stores, fence.proxy.async, mbarrier arrive.
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "qd_device_math.cuh"
#include "qd_internal.h"
#include "qd_tcfir.h"

namespace qd {

constexpr int kTcRows = 128;           // rows (of 64 samples) per MMA tile = UMMA M
constexpr int kTcRowSamples = 64;      // K = 128 reals = two 128-byte swizzle atoms of f16
constexpr int kTcEpiWarps = 8;           // two groups of four (TMEM lane quarter = warp & 3) taking alternate tiles
constexpr int kTcProdWarps = 8;
constexpr int kTcThreads = 32 * (kTcEpiWarps + 3 + kTcProdWarps);
constexpr uint32_t kTcStageBytes = 2 * kTcRows * 128; // one tile of A: two K atoms of 128 rows x 128 bytes
constexpr uint32_t kTcStages = 2;
constexpr uint32_t kTcRawBytes = kTcRows * kTcRowSamples * 2; // a tile's raw cs8 bytes: 16 KB, contiguous in the capture
constexpr uint32_t kTcRawMax = 8;
constexpr uint32_t kTcAnchors = 32; // ring of per-tile phase anchors, filled 16 tiles at a time

struct TcArgs {
    const uint8_t *src; // device pointer to raw sample src_base (absolute sample 0 sits on a 16-byte boundary)
    uint64_t src_base, src_end;
    uint64_t g0, g1; // outputs [g0, g1), absolute top-level indices
    float2 *out;     // out[g - g0]
    const uint8_t *bimg; // B operand image: [2 K atoms][N rows][128 bytes], swizzled
    uint32_t D, OPR, NOUT, NH, N, DMAX, XP;
    int32_t cown; // own output t of row b is g = OPR*b + cown + t
    int n_shift;
    double ratio[4];
    const double *sincos;
    SinCosK k;
    float s_hi, s_lo;
    int64_t row_first; // first row of tile 0
    uint32_t rows_eff; // rows a tile finalises: 128 - (DMAX - 1)
    uint32_t n_tiles, stages, acc_cols, tmem_cols;
    uint32_t off_b, off_x, off_a, off_raw; // byte offsets inside (1024-aligned) dynamic shared memory
    uint32_t raw_slots;                    // depth of the raw-byte ring the bulk copies fill
};

// ---------------------------------------------------------------------------- PTX
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait whose polls may be suspended by the hardware (time hint): a waiting role must not spend issue slots
__device__ __forceinline__ void mbar_wait_tc(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
                 "r"(parity), "r"(0x989680u)
                 : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint4 v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float *v)
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float *v)
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float2 lds_f2(uint32_t saddr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f2(uint32_t saddr, float2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr), "f"(v.x), "f"(v.y) : "memory");
}
// e^{i sum_s fl64(n ratio_s)} in f64: the product of the shifts' phasors at sample n, each phase formed as shift.rs:49 does
__device__ __forceinline__ double2 tc_phasor64(const TcArgs &a, int64_t n);
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand, 128-byte swizzle: 8-row groups of 128-byte rows, 1024 bytes apart (SBO); LBO unused; version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}

// Shared-memory header (at the 1024-aligned base): barriers and the TMEM base address
struct TcHeader {
    uint64_t full[4], empty[4], acc_full[2], acc_empty[2], raw_full[kTcRawMax], raw_empty[kTcRawMax], anc_full[kTcAnchors], anc_empty[kTcAnchors];
    uint32_t tmem_base;
    alignas(16) double2 anchor[kTcAnchors][2]; // per tile: e^{i ratio 64 * 128 q} for the two 128-row groups q the tile's rows lie in
    double2 w[kTcRows];            // e^{i ratio 64 r}, r < 128
};
static_assert(sizeof(TcHeader) <= 4096, "header region");
static_assert(kTcStages == 2, "a converter group owns one stage");

// four int8 I/Q bytes (two complex samples) -> two half2 (re, im): byte + 128 dropped into the mantissa of 1024.0h is
// the half 1024 + 128 + x; one exact packed subtraction leaves x
__device__ __forceinline__ void cvt_s8x4(uint32_t w, uint32_t &h0, uint32_t &h1)
{
    const uint32_t u = w ^ 0x80808080u;
    const uint32_t a = __byte_perm(u, 0x64646464u, 0x4140), b = __byte_perm(u, 0x64646464u, 0x4342);
    const __half2 k = __halves2half2(__ushort_as_half(0x6480), __ushort_as_half(0x6480)); // 1152.0
    const __half2 ra = __hsub2(*reinterpret_cast<const __half2 *>(&a), k), rb = __hsub2(*reinterpret_cast<const __half2 *>(&b), k);
    h0 = *reinterpret_cast<const uint32_t *>(&ra);
    h1 = *reinterpret_cast<const uint32_t *>(&rb);
}

__device__ __forceinline__ double2 tc_phasor64(const TcArgs &a, int64_t n)
{
    double2 w = make_double2(1.0, 0.0);
    for (int s = 0; s < a.n_shift; s++) {
        const double place = __dmul_rn(__ll2double_rn(n), a.ratio[s]);
        double cd, sd;
        sincos_f64k<false>(place, a.sincos, a.k, cd, sd);
        w = s == 0 ? make_double2(cd, sd) : make_double2(w.x * cd - w.y * sd, w.x * sd + w.y * cd);
    }
    return w;
}

// 16 raw bytes (8 samples from absolute sample n) of the capture, zero where the capture does not hold them
__device__ __forceinline__ uint4 tc_load_chunk(const TcArgs &a, int64_t n)
{
    if (n >= static_cast<int64_t>(a.src_base) && n + 8 <= static_cast<int64_t>(a.src_end))
        return ldg_stream_v4(a.src + (n - static_cast<int64_t>(a.src_base)) * 2);
    uint32_t w[4] = {0, 0, 0, 0};
    if (n + 8 > static_cast<int64_t>(a.src_base) && n < static_cast<int64_t>(a.src_end)) {
        for (int i = 0; i < 8; i++) {
            const int64_t m = n + i;
            if (m >= static_cast<int64_t>(a.src_base) && m < static_cast<int64_t>(a.src_end)) {
                const uint32_t v = __ldg(reinterpret_cast<const unsigned short *>(a.src) + (m - static_cast<int64_t>(a.src_base)));
                w[i >> 1] |= v << (16 * (i & 1));
            }
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// OPR_/NOUT_ != 0: the row geometry is known at compile time (the reference's default filter, args.rs:165: 40 taps,
// at decimate 8), so the epilogue keeps a row's partials in registers and is fully unrolled; 0/0: any geometry, the
// partials go through shared memory.
template <int OPR_, int NOUT_>
__global__ void __launch_bounds__(kTcThreads, 1) fk_tcfir(const __grid_constant__ TcArgs a)
{
    constexpr bool FIXED = OPR_ != 0;
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    TcHeader *hd = reinterpret_cast<TcHeader *>(smem);
    uint8_t *sB = smem + a.off_b;
    float *sX = reinterpret_cast<float *>(smem + a.off_x);
    uint8_t *sA = smem + a.off_a;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- setup: barriers, TMEM, the B image
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kTcStages; s++) {
            mbar_init(&hd->full[s], kTcProdWarps / 2);
            mbar_init(&hd->empty[s], 1);
        }
        for (uint32_t s = 0; s < kTcRawMax; s++) {
            mbar_init(&hd->raw_full[s], 1);
            mbar_init(&hd->raw_empty[s], kTcProdWarps / 2);
        }
        for (uint32_t s = 0; s < kTcAnchors; s++) {
            mbar_init(&hd->anc_full[s], 1);
            mbar_init(&hd->anc_empty[s], kTcEpiWarps / 2);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(&hd->acc_full[s], 1);
            mbar_init(&hd->acc_empty[s], kTcEpiWarps / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcEpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hd->tmem_base)), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const uint32_t nb = 2 * a.N * 128 / 16;
        const uint4 *g = reinterpret_cast<const uint4 *>(a.bimg);
        uint4 *s = reinterpret_cast<uint4 *>(sB);
        for (uint32_t i = threadIdx.x; i < nb; i += kTcThreads) s[i] = __ldg(g + i);
        fence_proxy_async();
    }
    if (threadIdx.x < kTcRows) hd->w[threadIdx.x] = tc_phasor64(a, static_cast<int64_t>(threadIdx.x) * kTcRowSamples);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hd->tmem_base;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint32_t n_my = blockIdx.x < a.n_tiles ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp < kTcEpiWarps) {
        // ================= epilogue: TMEM -> partials -> row rotation -> sum over the rows of an output -> global
        // group `grp` takes the CTA's tiles grp, grp + 2, ...: accumulator stage = grp
        const int grp = warp >> 2, quarter = warp & 3;
        const int r = quarter * 32 + lane; // tile row = TMEM lane
        const uint32_t opr = FIXED ? OPR_ : a.OPR, nout = FIXED ? NOUT_ : a.NOUT, nh = FIXED ? (2 * NOUT_ + 7) / 8 * 8 : a.NH;
        // exchange of partials between the rows of a tile: FIXED keeps a row's own partials in registers and double
        // buffers what its neighbours need; the general path passes everything and closes a tile with a second barrier
        const uint32_t xp = a.XP; // (general path; FIXED addresses its own layout)
        for (uint32_t k = grp; k < n_my; k += 2) {
            const uint32_t tile = blockIdx.x + k * gridDim.x;
            const uint32_t acc = grp, aph = (k >> 1) & 1;
            const int64_t b = a.row_first + static_cast<int64_t>(tile) * a.rows_eff + r;
            // The row phasor e^{i ratio 64 b} as a pure function of the absolute row: (anchor of the row's 128-row group,
            // computed for this tile by the bulk-copy warp) x (e^{i ratio 64 (b mod 128)}, tabulated at kernel start),
            // one f64 complex product rounded to f32.  s_hi rides along.
            float2 rot;
            {
                const int64_t b0 = b - r;
                mbar_wait_tc(&hd->anc_full[k % kTcAnchors], (k / kTcAnchors) & 1);
                const double2 an = hd->anchor[k % kTcAnchors][(b >> 7) - (b0 >> 7)], w = hd->w[b & 127];
                rot = make_float2(static_cast<float>(an.x * w.x - an.y * w.y) * a.s_hi, static_cast<float>(an.x * w.y + an.y * w.x) * a.s_hi);
                __syncwarp();
                if (lane == 0) mbar_arrive(&hd->anc_empty[k % kTcAnchors]);
            }
            float *xr = sX + static_cast<size_t>(grp) * kTcRows * xp + static_cast<size_t>(r) * xp;
            const uint32_t t0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * a.acc_cols;
            const int64_t gbase = static_cast<int64_t>(opr) * b + a.cown;
            float2 *o = a.out + (gbase - static_cast<int64_t>(a.g0));
            // every own output of every finalising row of the tile lies inside [g0, g1)?
            const int64_t g_lo = static_cast<int64_t>(opr) * (b - r) + a.cown, g_hi = g_lo + static_cast<int64_t>(opr) * a.rows_eff;
            const bool inside = g_lo >= static_cast<int64_t>(a.g0) && g_hi <= static_cast<int64_t>(a.g1);
            mbar_wait_tc(&hd->acc_full[acc], aph);
            tc_fence_after();
            if constexpr (FIXED) {
                constexpr int NC = 2 * NOUT_, NH = (2 * NOUT_ + 7) / 8 * 8, XPF = 2 * ((NOUT_ - OPR_) | 1); // 64-bit accesses at a pitch of 2 * odd words: no bank conflicts
                float part[NC];
                if constexpr (NH % 16 == 0) {
#pragma unroll
                    for (int c = 0; c < NH; c += 16) {
                        float hi[16], lo[16];
                        tc_ld16(t0 + c, hi);
                        tc_ld16(t0 + NH + c, lo);
                        tc_wait_ld();
#pragma unroll
                        for (int q = 0; q < 16; q++)
                            if (c + q < NC) part[c + q] = fmaf(lo[q], 0x1p-11f, hi[q]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NH; c += 8) {
                        float hi[8], lo[8];
                        tc_ld8(t0 + c, hi);
                        tc_ld8(t0 + NH + c, lo);
                        tc_wait_ld();
#pragma unroll
                        for (int q = 0; q < 8; q++)
                            if (c + q < NC) part[c + q] = fmaf(lo[q], 0x1p-11f, hi[q]);
                    }
                }
                tc_fence_before(); // the accumulator may be overwritten by the tile after next
                __syncwarp();
                if (lane == 0) mbar_arrive(&hd->acc_empty[acc]);
#pragma unroll
                for (int i = 0; i < NOUT_; i++) {
                    const float pr = part[2 * i], pi = part[2 * i + 1];
                    part[2 * i] = pr * rot.x - pi * rot.y;
                    part[2 * i + 1] = pr * rot.y + pi * rot.x;
                }
                // what the rows below need of this row: its partials for outputs that start in earlier rows
                const uint32_t xs = smem_u32(sX) + ((2 * grp + ((k >> 1) & 1)) * kTcRows + r) * (XPF * 4);
#pragma unroll
                for (int i = 0; i < NOUT_ - OPR_; i++) sts_f2(xs + 8 * i, make_float2(part[2 * i], part[2 * i + 1]));
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                if (static_cast<uint32_t>(r) < a.rows_eff) {
                    float2 y[OPR_];
#pragma unroll
                    for (int t = 0; t < OPR_; t++) {
                        y[t] = make_float2(part[2 * (NOUT_ - OPR_ + t)], part[2 * (NOUT_ - OPR_ + t) + 1]);
#pragma unroll
                        for (int d = 1; d * OPR_ < NOUT_; d++) {
                            const int ip = NOUT_ - OPR_ * (d + 1) + t;
                            if (ip >= 0) {
                                const float2 v = lds_f2(xs + d * (XPF * 4) + 8 * ip);
                                y[t].x += v.x;
                                y[t].y += v.y;
                            }
                        }
                    }
                    if (inside) {
                        if ((reinterpret_cast<uintptr_t>(o) & 15) == 0 && OPR_ % 2 == 0) {
#pragma unroll
                            for (int t = 0; t < OPR_; t += 2)
                                *reinterpret_cast<float4 *>(o + t) = make_float4(y[t].x, y[t].y, y[t + 1].x, y[t + 1].y);
                        } else {
#pragma unroll
                            for (int t = 0; t < OPR_; t++) o[t] = y[t];
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < OPR_; t++)
                            if (gbase + t >= static_cast<int64_t>(a.g0) && gbase + t < static_cast<int64_t>(a.g1)) o[t] = y[t];
                    }
                }
            } else {
                for (uint32_t c = 0; c < nh; c += 8) {
                    float hi[8], lo[8];
                    tc_ld8(t0 + c, hi);
                    tc_ld8(t0 + nh + c, lo);
                    tc_wait_ld();
#pragma unroll
                    for (int q = 0; q < 8; q += 2) {
                        const float pr = fmaf(lo[q], 0x1p-11f, hi[q]), pi = fmaf(lo[q + 1], 0x1p-11f, hi[q + 1]);
                        if (c + q < 2 * nout) {
                            xr[c + q] = pr * rot.x - pi * rot.y;
                            xr[c + q + 1] = pr * rot.y + pi * rot.x;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&hd->acc_empty[acc]);
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                if (static_cast<uint32_t>(r) < a.rows_eff) {
                    for (uint32_t t = 0; t < opr; t++) {
                        if (!inside && (gbase + t < static_cast<int64_t>(a.g0) || gbase + t >= static_cast<int64_t>(a.g1))) continue;
                        float yr = 0.0f, yi = 0.0f;
                        for (uint32_t d = 0; d < a.DMAX; d++) {
                            const int32_t ip = static_cast<int32_t>(nout) - static_cast<int32_t>(opr * (d + 1)) + static_cast<int32_t>(t);
                            if (ip < 0) break;
                            const float *p = xr + d * xp + 2 * ip;
                            yr += p[0];
                            yi += p[1];
                        }
                        o[t] = make_float2(yr, yi);
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); // the exchange buffer is free again
            }
        }
    } else if (warp == kTcEpiWarps) {
        // ================= MMA issue (one thread)
        const uint32_t idesc = (1u << 4) | ((a.N >> 3) << 17) | ((kTcRows >> 4) << 24); // f16 x f16 -> f32, K-major A and B
        const uint32_t b_base = smem_u32(sB);
        for (uint32_t k = 0; k < n_my; k++) {
            const uint32_t st = k % kTcStages, ph = (k / kTcStages) & 1, acc = k & 1, aph = (k >> 1) & 1;
            mbar_wait_tc(&hd->acc_empty[acc], aph ^ 1);
            mbar_wait_tc(&hd->full[st], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_base = smem_u32(sA + static_cast<size_t>(st) * kTcStageBytes);
                const uint32_t d_addr = tmem_base + acc * a.acc_cols;
#pragma unroll
                for (uint32_t ks = 0; ks < 8; ks++) { // K = 16 per instruction: four per 128-byte swizzle atom
                    const uint32_t ka = ks >> 2, k16 = ks & 3;
                    const uint64_t ad = umma_desc(a_base + ka * (kTcRows * 128) + k16 * 32);
                    const uint64_t bd = umma_desc(b_base + ka * (a.N * 128) + k16 * 32);
                    tc_mma_f16(d_addr, ad, bd, idesc, ks > 0 ? 1u : 0u);
                }
                tc_commit(&hd->empty[st]);    // the stage may be refilled once these MMAs have read it
                tc_commit(&hd->acc_full[acc]); // ... and the accumulator is complete
            }
            __syncwarp();
        }
    } else if (warp == kTcEpiWarps + 1) {
        // ================= bulk copies (one thread): a tile's 16 KB of raw bytes -> the raw ring (TMA, no registers)
        if (lane == 0) {
            uint32_t slot = 0, ph = 0;
            for (uint32_t k = 0; k < n_my; k++) {
                mbar_wait_tc(&hd->raw_empty[slot], ph ^ 1);
                const int64_t n0 = (a.row_first + static_cast<int64_t>(blockIdx.x + k * gridDim.x) * a.rows_eff) * kTcRowSamples;
                if (n0 >= static_cast<int64_t>(a.src_base) && n0 + kTcRows * kTcRowSamples <= static_cast<int64_t>(a.src_end)) {
                    mbar_expect_tx(&hd->raw_full[slot], kTcRawBytes);
                    bulk_g2s(smem + a.off_raw + slot * kTcRawBytes, a.src + (n0 - static_cast<int64_t>(a.src_base)) * 2, kTcRawBytes, &hd->raw_full[slot]);
                } else {
                    mbar_arrive(&hd->raw_full[slot]); // a tile at the edge of the capture: the converters read it themselves
                }
                if (++slot == a.raw_slots) slot = 0, ph ^= 1;
            }
        }
    } else if (warp == kTcEpiWarps + 2) {
        // ================= phase anchors: for every tile, e^{i ratio 64 * 128 q} (f64, the phase formed as shift.rs:49
        // does) of the two 128-row groups q its rows lie in, a ring ahead of the epilogue
        // (16 tiles per pass: lane l takes tile k0 + l/2, group l & 1)
        for (uint32_t k0 = 0; k0 < n_my; k0 += 16) {
            const uint32_t k = min(k0 + (lane >> 1), n_my - 1);
            const bool valid = k0 + (lane >> 1) < n_my;
            const int64_t b0 = a.row_first + static_cast<int64_t>(blockIdx.x + k * gridDim.x) * a.rows_eff;
            const double2 an = tc_phasor64(a, ((b0 >> 7) + (lane & 1)) * (128 * kTcRowSamples));
            mbar_wait_tc(&hd->anc_empty[k % kTcAnchors], ((k / kTcAnchors) & 1) ^ 1);
            if (valid) hd->anchor[k % kTcAnchors][lane & 1] = an;
            __syncwarp();
            if (valid && (lane & 1) == 0) mbar_arrive(&hd->anc_full[k % kTcAnchors]);
        }
    } else {
        // ================= converters: raw bytes -> f16 A tile (K-major, 128-byte swizzle); two groups of four warps
        // take alternate tiles (group = stage = k & 1), so two tiles are being converted at any time
        constexpr int kGroupWarps = kTcProdWarps / 2;
        constexpr int kPer = kTcRows * 8 / (32 * kGroupWarps); // 16-byte chunks per thread and tile
        const int cw = warp - kTcEpiWarps - 3, cg = cw / kGroupWarps, pw = cw % kGroupWarps;
        // A quarter warp (one 128-byte shared-memory wavefront) takes the first half of a row's raw bytes and the second
        // half of the next row's: 128 distinct bytes on the read side, and on the write side eight distinct 16-byte
        // columns of the swizzled layout (row parity flips the column parity).
        const uint32_t sub = (lane >> 2) & 1, row4 = 2 * (lane >> 4) + sub, ka = ((lane >> 3) & 1) ^ sub, cp = lane & 3;
        // chunk i of the thread: row = 4 * (i * kGroupWarps + pw) + row4, i.e. 16 * kGroupWarps / 4 rows further per i
        const uint32_t row_0 = 4 * pw + row4;
        constexpr uint32_t kStep = 4 * kGroupWarps * 128; // bytes from chunk i to chunk i + 1, raw and converted alike
        static_assert((4 * kGroupWarps) % 8 == 0, "the swizzle phase (row & 7) must not depend on i");
        const uint32_t g_off0 = row_0 * 128 + 16 * (ka * 4 + cp);
        const uint32_t s_off0 = ka * (kTcRows * 128) + row_0 * 128 + (((2 * cp) ^ (row_0 & 7)) << 4);
        const uint32_t s_off1 = ka * (kTcRows * 128) + row_0 * 128 + (((2 * cp + 1) ^ (row_0 & 7)) << 4);
        const uint32_t sA_u32 = smem_u32(sA), raw_u32 = smem_u32(smem + a.off_raw);
        uint32_t slot = cg % a.raw_slots, rph = 0;
        for (uint32_t k = cg; k < n_my; k += 2) {
            const uint32_t st = k % kTcStages, ph = (k / kTcStages) & 1;
            uint4 v[kPer];
            const int64_t n0 = (a.row_first + static_cast<int64_t>(blockIdx.x + k * gridDim.x) * a.rows_eff) * kTcRowSamples;
            const bool whole = n0 >= static_cast<int64_t>(a.src_base) && n0 + kTcRows * kTcRowSamples <= static_cast<int64_t>(a.src_end);
            mbar_wait_tc(&hd->raw_full[slot], rph);
            if (whole) {
#pragma unroll
                for (int i = 0; i < kPer; i++)
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                                 : "r"(raw_u32 + slot * kTcRawBytes + g_off0 + i * kStep)
                                 : "memory");
            } else {
#pragma unroll
                for (int i = 0; i < kPer; i++) v[i] = tc_load_chunk(a, n0 + ((g_off0 + i * kStep) >> 1));
            }
            mbar_wait_tc(&hd->empty[st], ph ^ 1);
            const uint32_t stage = sA_u32 + st * kTcStageBytes;
#pragma unroll
            for (int i = 0; i < kPer; i++) {
                uint4 lo4, hi4;
                cvt_s8x4(v[i].x, lo4.x, lo4.y);
                cvt_s8x4(v[i].y, lo4.z, lo4.w);
                cvt_s8x4(v[i].z, hi4.x, hi4.y);
                cvt_s8x4(v[i].w, hi4.z, hi4.w);
                sts_v4(stage + s_off0 + i * kStep, lo4);
                sts_v4(stage + s_off1 + i * kStep, hi4);
            }
            fence_proxy_async(); // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&hd->full[st]);
                mbar_arrive(&hd->raw_empty[slot]); // the slot's bytes have been read and used: it may be refilled
            }
            slot += 2; // the group's next tile is two tiles on
            if (slot >= a.raw_slots) slot -= a.raw_slots, rph ^= 1;
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == kTcEpiWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------- host side
static int64_t floor_div(int64_t a, int64_t b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

bool tcfir_geometry(uint32_t L, uint32_t D, TcGeom *g)
{
    if (D < 2 || D > 32 || (D & (D - 1)) || L < 1 || L > 4096) return false;
    const int64_t i0 = L - L / 2;
    g->OPR = kTcRowSamples / D;
    // outputs whose taps meet row 0 (samples [0, 64)): g*D + i0 <= 63 and g*D + i0 + L - 1 >= 0
    const int64_t gmax = floor_div(63 - i0, D), gmin = -floor_div(i0 + L - 1, D);
    g->c0 = static_cast<int32_t>(gmin);
    g->NOUT = static_cast<uint32_t>(gmax - gmin + 1);
    g->NH = (2 * g->NOUT + 7) / 8 * 8;
    g->N = 2 * g->NH;
    if (g->N % 16) g->NH += 8, g->N = 2 * g->NH; // UMMA with M = 128: N a multiple of 16
    if (g->N < 16 || g->N > 256) return false;
    g->DMAX = (g->NOUT + g->OPR - 1) / g->OPR;
    if (g->DMAX >= 64) return false;
    g->cown = g->c0 + static_cast<int32_t>(g->NOUT) - static_cast<int32_t>(g->OPR);
    g->XP = (2 * g->NOUT) | 1;
    // B image + exchange buffers + the A stages must fit one CTA's shared memory (tcfir_smem_layout)
    const bool fixed = g->OPR == 8 && g->NOUT == 13;
    const uint32_t x_bytes = fixed ? 4 * kTcRows * (2 * ((g->NOUT - g->OPR) | 1)) * 4 : 2 * kTcRows * g->XP * 4;
    const uint32_t off_a = (4096 + 2 * g->N * 128 + x_bytes + 1023) / 1024 * 1024;
    if (off_a + kTcStages * kTcStageBytes + 2 * kTcRawBytes > 227 * 1024 - 1024) return false;
    return true;
}

static uint16_t f16_bits(float v)
{
    return static_cast<__half_raw>(__float2half_rn(v)).x;
}
static float f16_value(uint16_t b)
{
    __half_raw r;
    r.x = b;
    return __half2float(__half(r));
}

// The B operand as it lies in shared memory; *s_hi, *s_lo: what the two column halves are multiplied by
void tcfir_b_image(const TcGeom &g, const float *taps, uint32_t L, uint32_t D, double ratio_sum, std::vector<uint8_t> &img, float *s_hi, float *s_lo)
{
    const int64_t i0 = L - L / 2;
    double gmaxabs = 0.0;
    for (uint32_t j = 0; j < L; j++) gmaxabs = std::max(gmaxabs, std::fabs(static_cast<double>(taps[j]) / 127.0));
    int e = 0;
    if (gmaxabs > 0.0) frexp(gmaxabs, &e); // gmaxabs = m * 2^e, m in [0.5, 1)
    const double sigma = ldexp(1.0, 13 - e); // |g sigma| < 2^13
    *s_hi = static_cast<float>(1.0 / sigma);
    *s_lo = static_cast<float>(ldexp(1.0 / sigma, -11));
    img.assign(static_cast<size_t>(2) * g.N * 128, 0);
    for (uint32_t n = 0; n < g.N; n++) {
        const uint32_t half = n / g.NH, cn = n % g.NH;
        if (cn >= 2 * g.NOUT) continue;
        const uint32_t ip = cn / 2, ri = cn % 2;
        for (uint32_t k = 0; k < 128; k++) {
            const int64_t kk = k / 2, c = k % 2;
            const int64_t j = kk - (static_cast<int64_t>(g.c0) + ip) * D - i0;
            if (j < 0 || j >= static_cast<int64_t>(L)) continue;
            const double f = static_cast<double>(taps[j]) / 127.0 * sigma, ph = ratio_sum * static_cast<double>(kk);
            const double gr = f * cos(ph), gi = f * sin(ph);
            // (xr + i xi)(gr + i gi): re = xr gr - xi gi, im = xr gi + xi gr
            const double v = ri == 0 ? (c == 0 ? gr : -gi) : (c == 0 ? gi : gr);
            const uint16_t hb = f16_bits(static_cast<float>(v));
            const uint16_t lb = f16_bits(static_cast<float>((v - static_cast<double>(f16_value(hb))) * 2048.0));
            const uint32_t kat = k / 64, e64 = k % 64;
            const size_t at = static_cast<size_t>(kat) * g.N * 128 + static_cast<size_t>(n) * 128 + (((e64 / 8) ^ (n & 7)) << 4) + (e64 % 8) * 2;
            const uint16_t bits = half == 0 ? hb : lb;
            memcpy(&img[at], &bits, 2);
        }
    }
}

int launch_tcfir(Chain &c, const TcGeom &g, const uint8_t *d_bimg, float s_hi, float s_lo, uint32_t L, uint32_t D, int n_shift,
                 const double *ratios, const uint8_t *d_src, uint64_t src_base, uint64_t src_end, uint64_t g0, uint64_t g1, float2 *d_out)
{
    if (g1 <= g0) return QD_OK;
    TcArgs a;
    memset(&a, 0, sizeof a);
    a.src = d_src, a.src_base = src_base, a.src_end = src_end;
    a.g0 = g0, a.g1 = g1, a.out = d_out, a.bimg = d_bimg;
    a.D = D, a.OPR = g.OPR, a.NOUT = g.NOUT, a.NH = g.NH, a.N = g.N, a.DMAX = g.DMAX, a.XP = g.XP, a.cown = g.cown;
    a.n_shift = n_shift;
    for (int i = 0; i < n_shift; i++) a.ratio[i] = ratios[i];
    a.sincos = c.ctx->d_sincos;
    a.k = make_sincos_k();
    a.s_hi = s_hi, a.s_lo = s_lo;
    // output g is finalised by the row that holds its first sample g*D + i0
    const int64_t i0 = L - L / 2;
    a.row_first = floor_div(static_cast<int64_t>(g0) * D + i0, kTcRowSamples);
    const int64_t row_last = floor_div(static_cast<int64_t>(g1 - 1) * D + i0, kTcRowSamples);
    a.rows_eff = kTcRows - (g.DMAX - 1);
    const uint64_t tiles = static_cast<uint64_t>(row_last - a.row_first) / a.rows_eff + 1;
    if (tiles >= (uint64_t(1) << 31)) return set_error(QD_E_INVALID_ARG, "internal: too many tensor-core tiles");
    a.n_tiles = static_cast<uint32_t>(tiles);
    a.acc_cols = g.N;
    a.tmem_cols = 32;
    while (a.tmem_cols < 2 * a.acc_cols) a.tmem_cols *= 2;
    a.off_b = 4096;
    a.off_x = a.off_b + 2 * g.N * 128;
    const bool fixed = g.OPR == 8 && g.NOUT == 13; // the reference's default filter (40 taps) at decimate 8: config 2
    const uint32_t x_bytes = fixed ? 4 * kTcRows * (2 * ((g.NOUT - g.OPR) | 1)) * 4 : 2 * kTcRows * g.XP * 4;
    a.off_a = (a.off_x + x_bytes + 1023) / 1024 * 1024;
    const uint32_t cap = 227 * 1024 - 1024;
    a.off_raw = a.off_a + kTcStages * kTcStageBytes;
    if (a.off_raw + 2 * kTcRawBytes > cap) return set_error(QD_E_INVALID_ARG, "internal: tensor-core FIR does not fit shared memory");
    a.stages = kTcStages;
    a.raw_slots = std::min<uint32_t>(kTcRawMax, (cap - a.off_raw) / kTcRawBytes);
    const size_t smem = 1024 + a.off_raw + static_cast<size_t>(a.raw_slots) * kTcRawBytes;
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(tiles, static_cast<uint64_t>(c.ctx->sm_count)));
    if (fixed) {
        QD_CUDA(cudaFuncSetAttribute(fk_tcfir<8, 13>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        fk_tcfir<8, 13><<<grid, kTcThreads, smem, c.stream>>>(a);
    } else {
        QD_CUDA(cudaFuncSetAttribute(fk_tcfir<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        fk_tcfir<0, 0><<<grid, kTcThreads, smem, c.stream>>>(a);
    }
    QD_LAUNCHED();
    return QD_OK;
}

} // namespace qd
