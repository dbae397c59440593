// qd_fir_kernel.cuh -- the fused hot-path kernel: decode + NCO mix + decimating FIR in one pass.
//
// Canonical chain: From(cs8|cu8|cs16|cf32) -> Shift* -> LowPass, feeding do_write chunks, read_at
// units or sparkfft windows (samples.rs:72-93 -> shift.rs:46-54 -> filter.rs:54-124).  Each raw sample
// crosses HBM once: a persistent CTA stages a tile of raw bytes plus its (taps-1) halo in shared
// memory with a 1-D bulk async copy (TMA, cp.async.bulk + mbarrier; the next tile's copy is in flight while
// the current tile is filtered), decodes and
// mixes it once into a polyphase shared-memory layout, then computes ONLY the kept outputs, each
// thread holding R consecutive outputs in registers so every staged sample is reused from registers.
//
// Arithmetic.  Per output the taps are applied in ascending order, so the reference's zero-truncated
// tail (filter.rs:107-124: taps past the end of the caller's raw buffer contribute nothing) is simply
// an earlier loop exit.  EXACT mode rounds the product and the sum separately (FMUL2 then FFMA2 by an
// opaque 1.0: ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into one FFMA2, which would change the
// rounding) and so is bit-identical to the reference's `acc += x * f`; FAST mode uses one FFMA2.
#pragma once
// Shared by qd_fast.cu (host side, fk_tail) and the per-decimation translation units qd_fast_dN.cu, which
// instantiate fk_fir for one decimation each so that they compile in parallel.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "qd_device_math.cuh"
#include "qd_internal.h"
#include "qd_stft_epilogue.cuh"

namespace qd {

#ifndef QD_EXP_EXACT_CTAS
#define QD_EXP_EXACT_CTAS 5
#endif
constexpr int kMaxTapPairs = 1024;
constexpr int kMaxLeadShifts = 4;

struct alignas(16) FirTaps {
    float2 t[kMaxTapPairs]; // (f[j], f[j]) pairs, zero padded to Q*D
    float s[kMaxTapPairs];  // the same taps once each: the run-time-length loop fetches four per constant load and
                            // lets the packed instructions broadcast them
};

struct FirArgs {
    const uint8_t *src; // device pointer to raw sample `src_base`
    uint64_t src_base;
    uint64_t src_end; // src_base + resident samples
    int fmt;
    int n_shift;
    double ratio[kMaxLeadShifts];
    const double *sincos;
    SinCosK k;
    uint32_t L;          // filter length
    uint64_t off0;       // top-level (decimated) index of unit 0's first output
    uint64_t n_call;     // outputs per unit (the n of LowPass::read_at)
    uint64_t S;          // unit stride in top-level samples
    uint64_t n_units;
    int contiguous; // S == n_call: the units tile the output stream
    int no_carry;   // experiments: 1 = every tile decodes its whole span (no overlap carried between a CTA's consecutive tiles)
    uint32_t tiles_per_unit;
    uint64_t n_tiles;
    uint64_t tile_first; // contiguous mode: tiles are numbered from ABSOLUTE top-level output 0 (tile A holds outputs
                         // [A*T, (A+1)*T)); the launch covers tiles tile_first .. tile_first + n_tiles - 1, the first and
                         // last of them partly.  A tile's phase anchor, lane phasors and recurrences therefore depend
                         // only on absolute position, never on where a launch (host segment, shard) starts.
    uint64_t total_out; // contiguous mode: outputs to compute (<= n_units * n_call)
    uint32_t raw_cap; // bytes per raw staging buffer
    float2 *out;      // [n_units][n_call]
    float2 one;       // (1, 1), opaque to ptxas
    // FAST mode NCO: all leading shifts merged into one rotation of ratio_sum per sample
    float2 rot[4];    // e^{i k ratio_sum}, k = 0..3 (k = 0 unused)
    float2 rot_step;  // e^{i 4*NT ratio_sum}: from one group of a thread to its next
    float rot2c;      // 2 cos(ratio_sum)
    // lean cs8 path (one shift): |ratio[0]| = rmant * 2^rexp exactly, rsign = +-1 (0: ratio is zero)
    uint64_t rmant;
    int rexp;
    int rsign;
    int ncall_log2;   // n_call = 2^ncall_log2, or -1
    // Truncated window tails as SNAPSHOTS of the running sums (stream mode feeding overlapping sparkfft windows).
    // The reference applies taps in ascending order (filter.rs:113-120), so the zero-truncated output k of the read
    // (u*S, W) -- J = (W - k)*D + L/2 < L taps -- is exactly the stream output's accumulator after its first J taps.
    // Stream output q (relative to off0) belongs to the tail of window u = (q - (W - T)) / S at place
    // rr = (q - (W - T)) % S when rr < T; with S >= T that is at most one window per output.  tail_out[u*T + rr].
    float2 *tail_out; // nullptr: no snapshots
    uint32_t tail_W, tail_S, tail_T;
    uint64_t tail_units;
    // Fused sparkfft sink (fft.W != 0): the units of the launch are back-to-back windows of fft.W outputs (stride =
    // width); a tile holds whole windows, so the kernel transforms them in shared memory right after the filter and
    // writes glyph indices (and magnitudes) instead of the cf32 outputs -- the decimated stream never reaches HBM.
    FftArgs fft;
    // fft.W != 0 in STREAM mode (overlapping windows cut from the stream, FUSE = 2): each CTA takes a contiguous run
    // of tiles and carries the last W - 1 outputs (and their snapshots) from tile to tile in shared memory, so every
    // window is transformed by the tile that holds its last output.  fuse_S = the window stride in outputs.
    uint32_t fuse_S;
    uint32_t carry_off; // byte offset of the carry buffers inside dynamic shared memory
    // A SECOND lowpass between this launch's stream and the windows (st2_L != 0; config 5's `lowpass | lowpass |
    // sparkfft`): its output k (st2_L taps at decimation st2_D over this launch's outputs, none of them truncated)
    // is computed by the tile that holds its last input, from the carried outputs, and the windows are cut from
    // those second-stage outputs.  Neither intermediate stream reaches HBM.
    uint32_t st2_L, st2_D;
    uint64_t st2_total; // second-stage outputs the windows need
    float st2_taps[256];
};

// Per-tile phase state of the lean FAST decode, computed by one thread while the previous tile is filtered
struct LeanPhase {
    double ac, as;  // e^{i n_tile0 ratio} with the product taken exactly
    uint64_t m64k;  // (rmant << (64 - k)) mod 2^64: n * m64k mod 2^64 = the k discarded bits of n*rmant, left aligned
    uint32_t mk32;  // its high word: the per-sample increment of the 32-bit fraction
    float esc;      // rsign * 2^(rexp + k - 32): fraction -> radians
    int ok;         // the tile lies inside one binade of n*ratio (k is constant)
    int pad;
};
constexpr int kSmemHeader = 64; // two mbarriers + LeanPhase

// ---------------------------------------------------------------------------- PTX helpers
template <bool EXACT>
__device__ __forceinline__ float2 mac(float2 acc, float2 x, float2 tap, float2 one)
{
    if (EXACT) return fma2(mul2(x, tap), one, acc); // fl(acc + fl(x*f)): filter.rs:119
    return fma2(x, tap, acc);
}

// ---------------------------------------------------------------------------- exact integer decode
// the same, on an (I, Q) pair at once
__device__ __forceinline__ float2 div_exact2(float2 x, float den, float c)
{
    const float2 c2 = make_float2(c, c);
    const float2 q0 = mul2(x, c2);
    const float2 r = fma2(q0, make_float2(-den, -den), x);
    return fma2(r, c2, q0);
}

// cu8 / cs16 (lib.rs:252-253): fl(fl(x / den) - off) = fl(x * fl(1/den) - off) with ONE rounding, for every u8 / 255 and
// every i16 / 65535 (exhaustive, exact rational arithmetic: tests/test_decode_trick.py) -- the result's ulp is so
// much coarser than the quotient's that the first rounding never matters.  One packed FMA instead of four operations.
__device__ __forceinline__ float2 dec_fused2(float2 x, float c, float off) { return fma2(x, make_float2(c, c), make_float2(off, off)); }

// One sample of a 4-sample group held in w[] (FMT is compile time here)
// SCALED = false (FAST mode, cs8): the integer value itself; the kernel's taps carry the 1/127
template <int FMT, bool SCALED = true>
__device__ __forceinline__ float2 decode_in_group(const uint32_t (&w)[8], int i, float2 one)
{
    if (FMT == QD_FMT_CF32) return make_float2(__uint_as_float(w[2 * i]), __uint_as_float(w[2 * i + 1])); // bit copy
    if (FMT == QD_FMT_CS8) { // lib.rs:251
        const uint32_t h = w[i >> 1] >> ((i & 1) * 16);
        const float x = static_cast<float>(static_cast<int>(static_cast<signed char>(h & 0xff)));
        const float y = static_cast<float>(static_cast<int>(static_cast<signed char>((h >> 8) & 0xff)));
        if (!SCALED) return make_float2(x, y);
        return div_exact2(make_float2(x, y), 127.0f, 1.0f / 127.0f);
    }
    if (FMT == QD_FMT_CU8) { // lib.rs:252: x/255 - 127.5 (the subtraction as q*1 + (-127.5), one rounding)
        const uint32_t h = w[i >> 1] >> ((i & 1) * 16);
        const float x = static_cast<float>(h & 0xff), y = static_cast<float>((h >> 8) & 0xff);
        return dec_fused2(make_float2(x, y), 1.0f / 255.0f, -127.5f);
    }
    // cs16, lib.rs:253
    const float x = static_cast<float>(static_cast<int>(static_cast<short>(w[i] & 0xffff)));
    const float y = static_cast<float>(static_cast<int>(static_cast<short>(w[i] >> 16)));
    return dec_fused2(make_float2(x, y), 1.0f / 65535.0f, -32767.5f);
}

// ---------------------------------------------------------------------------- tile geometry
constexpr int pitch_for(int G, int cols)
{
    // elements are 16 bytes (a PAIR of consecutive samples): a quarter-warp storing 8 of them must hit 8
    // distinct 16-byte bank groups.  With G >= 8 consecutive lanes store to consecutive rows (odd pitch);
    // with G = 4 they cover 4 rows x 2 columns (pitch = 2 mod 8)
    if (G >= 8) return cols | 1;
    int p = cols;
    while (p % 8 != 8 / G) p++;
    return p;
}

// Outputs per tile.  R*NT by default; when the filter length is known at compile time the tile is trimmed so that
// its raw span, (T-1)*D + L samples, is a whole number of decode iterations (4*NT samples each): no nearly empty
// last iteration for one warp to run while the others wait at the barrier.  Kept a multiple of R (a thread's
// outputs never straddle a unit) and even (16-byte aligned output rows).
constexpr int tile_outputs(int D, int R, int NT, int LS)
{
    const int t_out = R * NT;
    if (LS <= 0) return t_out;
    const int iters = ((t_out - 1) * D + LS) / (4 * NT);
    if (iters < 1 || iters * 4 * NT < LS + D) return t_out;
    int t = (iters * 4 * NT - LS) / D + 1;
    if (t > t_out) t = t_out;
    t -= t % R;
    t -= t % 2;
    return t >= t_out / 2 ? t : t_out;
}

template <int D, int R, int NT, int LMAX = kMaxTapPairs>
struct FirGeom {
    static constexpr int DR = D * R; // polyphase period
    static constexpr int G = DR / 4; // physical rows are grouped by (row & 3)
    static constexpr int LOG_DR = (DR == 16) ? 4 : (DR == 32) ? 5 : 6;
    static constexpr int LOG_G = LOG_DR - 2;
    static constexpr int T_OUT = R * NT;
    static constexpr int T_TILE = tile_outputs(D, R, NT, LMAX == kMaxTapPairs ? 0 : LMAX); // outputs a tile really holds
    // samples the tile's threads touch: whole tap blocks of D for the longest filter this layout holds, plus
    // the tail of a partial last decode group where one can exist
    static constexpr int SPAN = (T_OUT - 1) * D + (LMAX + D - 1) / D * D + ((LMAX % 4 || D % 4) ? 3 : 0);
    static constexpr int COLS = (SPAN + DR - 1) / DR;
    static constexpr int PITCH = pitch_for(G, COLS); // float4 (sample pairs) per row; there are DR/2 rows
    static constexpr size_t X_BYTES = static_cast<size_t>(DR / 2) * PITCH * sizeof(float4);
    static_assert(DR == 16 || DR == 32 || DR == 64, "polyphase period must be 16, 32 or 64");
};

struct TileGeo {
    uint64_t n_tile0; // absolute raw index of local sample 0 (first tap of the tile's first output slot)
    int64_t f0;       // flat output index of the tile's first slot relative to the launch's first output: negative
                      // for the launch's first tile when the launch starts inside it (contiguous mode)
    uint64_t unit;    // non-contiguous mode
    uint64_t u0;      // contiguous mode: unit that holds the tile's first output
    uint32_t cnt;     // output slots of this tile up to its last wanted output
    uint32_t skip;    // leading slots that belong to an earlier launch (only the first tile of a launch)
};

template <int D, int T_OUT>
__device__ __forceinline__ TileGeo tile_geo(const FirArgs &a, uint64_t tile)
{
    TileGeo g;
    const uint32_t i0 = a.L - a.L / 2; // convoluted[L + k*D] is loop index L + k*D - L/2 (filter.rs:78-80,111)
    if (a.contiguous) {
        const uint64_t first = (a.tile_first + tile) * T_OUT; // absolute top-level index of the tile's first slot
        g.skip = first < a.off0 ? static_cast<uint32_t>(a.off0 - first) : 0u;
        g.f0 = static_cast<int64_t>(first - a.off0);
        g.cnt = static_cast<uint32_t>(min(static_cast<uint64_t>(T_OUT), a.off0 + a.total_out - first));
        g.unit = 0;
        const uint64_t fpos = g.skip ? 0 : static_cast<uint64_t>(g.f0);
        g.u0 = a.ncall_log2 >= 0 ? (fpos >> a.ncall_log2) : fpos / a.n_call;
        g.n_tile0 = first * D + i0;
    } else {
        if (a.n_tiles <= 0xffffffffull) { // 32-bit division
            const uint32_t u = static_cast<uint32_t>(tile) / a.tiles_per_unit;
            g.unit = u;
        } else {
            g.unit = tile / a.tiles_per_unit;
        }
        const uint32_t k0 = static_cast<uint32_t>(tile - g.unit * a.tiles_per_unit) * T_OUT;
        g.cnt = static_cast<uint32_t>(min(static_cast<uint64_t>(T_OUT), a.n_call - k0));
        g.skip = 0;
        g.f0 = static_cast<int64_t>(g.unit * a.n_call + k0);
        g.u0 = 0;
        g.n_tile0 = (a.off0 + g.unit * a.S + k0) * D + i0;
    }
    return g;
}

// ---------------------------------------------------------------------------- decode + mix stage
// Samples are kept in PAIRS (16 bytes): pair P = l / 2 lives at X[prow(P mod DR/2)][P div DR/2] with
// prow(r) = (r & 1) * G + (r >> 1), G = DR/4.  A lane handles one 16-byte-aligned group of 4 raw samples
// (two pairs); its neighbours handle the next groups, so for a fixed pair of the group a quarter-warp
// writes to consecutive physical rows: conflict-free 128-bit stores.  The FIR reads one row at consecutive
// columns with 128-bit loads.  ALIGNED: the tile's first sample sits on a group boundary (lead % 4 == 0),
// the common case, and the two stores of a group are one base address plus compile-time offsets.
// FAST mode complex multiply: contraction allowed
__device__ __forceinline__ float2 cmul_fast(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

template <bool HALF>
__device__ __forceinline__ float2 mix_exact(float2 v, double nd, double ratio, const FirArgs &a)
{
    double c, sn;
    sincos_f64k<HALF>(__dmul_rn(nd, ratio), a.sincos, a.k, c, sn); // place = (off + i) as f64 * ratio, shift.rs:49
    return cmul_exact(v, make_float2(static_cast<float>(c), static_cast<float>(sn)));
}

// l_lo: local samples below it belong to an earlier launch (first tile of a launch only): they are neither
// loaded nor stored, but every recurrence steps through their groups as if they were, so that the samples that
// are wanted come out exactly as in a launch that covers the whole tile.
template <int FMT, int D, int R, int NT, int LMAX, bool ALIGNED, bool FASTMIX>
__device__ __forceinline__ void decode_tile(const FirArgs &a, const uint8_t *raw, uint32_t lead, uint32_t l_lo, uint32_t n_dec,
                                            uint64_t n_tile0, float2 *__restrict__ X, int tid)
{
    using Gm = FirGeom<D, R, NT, LMAX>;
    const int n_have = static_cast<int>(n_dec + lead);
    const int n_skip = static_cast<int>(l_lo + lead); // raw-group-relative index of the first wanted sample
    const uint32_t n_groups = static_cast<uint32_t>(n_have + 3) / 4;
    // absolute index of raw group 0, sample 0, as an exact f64 (indices stay far below 2^53)
    const double base_d = __ull2double_rn(n_tile0 - lead);
    const int n_shift = a.n_shift;
    const float2 one = a.one;
    // FAST: the phase is evaluated in f64 exactly as the reference does it (shift.rs:49) once per thread and
    // tile, at the thread's first sample, and carried forward by f32 rotations (<= 8 steps of 4*NT samples)
    float2 ph_g = make_float2(1.0f, 0.0f);
    if (FASTMIX && n_shift) {
        // anchor = e^{i * (exact n*ratio)}: the f64 phase p = fl64(n*ratio) is off the exact product by
        // delta = -fma(n, ratio, -p), removed here and re-applied per sample below
        const double nd = __dadd_rn(base_d, static_cast<double>(4 * tid));
        for (int s = 0; s < n_shift; s++) {
            const double p = __dmul_rn(nd, a.ratio[s]);
            const float e = static_cast<float>(fma(nd, a.ratio[s], -p));
            double c, sn;
            sincos_f64k(p, a.sincos, a.k, c, sn);
            const float cf = static_cast<float>(c), sf = static_cast<float>(sn);
            ph_g = cmul_fast(ph_g, make_float2(fmaf(-e, sf, cf), fmaf(e, cf, sf)));
        }
    }
    // cf32 groups come from global memory (L2, after the bulk prefetch): keep the next group's loads in flight
    uint4 nlo = make_uint4(0, 0, 0, 0), nhi = make_uint4(0, 0, 0, 0);
    auto fetch = [&](uint32_t grp) {
        nlo = nhi = make_uint4(0, 0, 0, 0);
        if (static_cast<int>(4 * grp + 1) >= n_skip) nlo = ldg_stream_v4(reinterpret_cast<const uint4 *>(raw) + 2 * grp);
        if (static_cast<int>(4 * grp + 2) < n_have && static_cast<int>(4 * grp + 3) >= n_skip)
            nhi = ldg_stream_v4(reinterpret_cast<const uint4 *>(raw) + 2 * grp + 1);
    };
    if (FMT == QD_FMT_CF32 && static_cast<uint32_t>(tid) < n_groups) fetch(tid);
    for (uint32_t grp = tid; grp < n_groups; grp += NT) {
        uint32_t w[8];
        if (FMT == QD_FMT_CF32) {
            const uint4 lo = nlo, hi = nhi;
            if (grp + NT < n_groups) fetch(grp + NT);
            w[0] = lo.x, w[1] = lo.y, w[2] = lo.z, w[3] = lo.w, w[4] = hi.x, w[5] = hi.y, w[6] = hi.z, w[7] = hi.w;
        }
        const int g4 = static_cast<int>(4 * grp);
        if (g4 + 3 < n_skip) { // nothing wanted in this group: only the phasor recurrence moves on
            if (FASTMIX && n_shift) ph_g = cmul_fast(ph_g, a.rot_step);
            continue;
        }
        // integer formats: shared memory when the tile is staged, else global memory (never below the first wanted group)
        if (FMT == QD_FMT_CS16) {
            const uint4 v = *(reinterpret_cast<const uint4 *>(raw) + grp);
            w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
        } else if (FMT != QD_FMT_CF32) {
            const uint2 v = *(reinterpret_cast<const uint2 *>(raw) + grp);
            w[0] = v.x, w[1] = v.y;
        }
        const bool interior = g4 >= n_skip && g4 + 3 < n_have;
        float2 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = decode_in_group<FMT, !FASTMIX>(w, i, one);
        if (FASTMIX) {
            if (n_shift) {
                // the reference's phase is fl64(n*ratio) (shift.rs:49): its rounding error against the exact
                // product, e = n*ratio - p, is recovered exactly by one FMA and applied as a small rotation
                const double nd0 = __dadd_rn(base_d, static_cast<double>(g4));
                const double r0 = a.ratio[0];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double nd = __dadd_rn(nd0, static_cast<double>(i));
                    float e = static_cast<float>(fma(nd, r0, -__dmul_rn(nd, r0)));
                    if (n_shift > 1)
                        for (int s = 1; s < n_shift; s++) e += static_cast<float>(fma(nd, a.ratio[s], -__dmul_rn(nd, a.ratio[s])));
                    float2 ph = i == 0 ? ph_g : cmul_fast(ph_g, a.rot[i]);
                    ph = make_float2(fmaf(e, ph.y, ph.x), fmaf(-e, ph.x, ph.y)); // * (1 - i e) = * e^{i delta}
                    v[i] = cmul_fast(v[i], ph);
                }
                ph_g = cmul_fast(ph_g, a.rot_step);
            }
        } else if (n_shift) {
            const double nd0 = __dadd_rn(base_d, static_cast<double>(g4));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double nd = __dadd_rn(nd0, static_cast<double>(i));
                v[i] = mix_exact<(D <= 8)>(v[i], nd, a.ratio[0], a);
                for (int s = 1; s < n_shift; s++) v[i] = mix_exact<(D <= 8)>(v[i], nd, a.ratio[s], a);
            }
        }
        float4 *X4 = reinterpret_cast<float4 *>(X);
        if (ALIGNED) {
            const uint32_t gc = grp - (lead >> 2); // local group index (wraps for the skipped lead groups)
            float4 *xb = X4 + (gc & (Gm::G - 1)) * Gm::PITCH + (gc >> Gm::LOG_G);
            if (interior) {
                xb[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
                xb[Gm::G * Gm::PITCH] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (g4 + i >= n_skip && g4 + i < n_have)
                        reinterpret_cast<float2 *>(xb + (i >> 1) * Gm::G * Gm::PITCH)[i & 1] = v[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int l = g4 + i - static_cast<int>(lead);
                if (l < static_cast<int>(l_lo) || l >= static_cast<int>(n_dec)) continue;
                const uint32_t pr = (static_cast<uint32_t>(l) >> 1) & (Gm::DR / 2 - 1);
                float4 *el = X4 + ((pr & 1) * Gm::G + (pr >> 1)) * Gm::PITCH + (static_cast<uint32_t>(l) >> Gm::LOG_DR);
                reinterpret_cast<float2 *>(el)[l & 1] = v[i];
            }
        }
    }
}

// ---------------------------------------------------------------------------- lean FAST decode (cs8, <= 1 shift)
// The common FAST case gets its own loop.  Decode: a byte b = x + 128 dropped into the mantissa of 2^23 is
// the float 2^23 + b, and subtracting 2^23 + 128 leaves x exactly (two PRMT and one packed add per sample, no
// integer->float conversion).  Phase: thread t's first group starts 4t samples into the tile, so its phasor
// is (tile anchor) * e^{i 4t ratio}; the anchor is evaluated in f64 once per tile by one thread, the second
// factor once per kernel.  The reference's phase is fl64(n * ratio) (shift.rs:49), off the exact product by
// the k bits the multiply rounds away.  Those bits are n * rmant mod 2^k: kept left-aligned in a 64-bit
// integer they advance by one wrapping add per group, and their top word read as a signed fraction of an
// ulp is the rounding error (ties excepted: they round to even, here always up).
__device__ __forceinline__ void lean_phase(const FirArgs &a, uint64_t n0, uint32_t span, LeanPhase *ph)
{
    // always from scratch (not stepped from the CTA's previous tile): a tile's result must not depend on which
    // launch -- host-path segment, shard -- it is part of
    const double nd = __ull2double_rn(n0), r = a.ratio[0];
    const double p = __dmul_rn(nd, r);
    const double e = fma(nd, r, -p); // exact product minus the rounded one
    double c, s;
    sincos_f64k(p, a.sincos, a.k, c, s);
    ph->ac = fma(-e, s, c);
    ph->as = fma(e, c, s);
    const uint64_t M = a.rmant;
    auto bitlen = [&](uint64_t n) {
        const uint64_t hi = __umul64hi(n, M), lo = n * M;
        return hi ? 128 - __clzll(static_cast<long long>(hi)) : 64 - __clzll(static_cast<long long>(lo));
    };
    const int b0 = bitlen(n0), b1 = bitlen(n0 + span);
    const int k = b0 - 53; // bits rounded away (<= 64 since rmant < 2^53)
    const int ee = a.rexp + k - 32;
    uint64_t m64k = 0;
    float esc = 0.0f;
    if (k >= 1 && a.rsign != 0) {
        m64k = k >= 64 ? M : (M << (64 - k));
        if (ee >= -126 && ee <= 127) esc = __int_as_float((127 + ee) << 23) * static_cast<float>(a.rsign);
    }
    ph->m64k = m64k;
    ph->mk32 = static_cast<uint32_t>(m64k >> 32);
    ph->esc = esc;
    ph->ok = (b0 == b1) ? 1 : 0;
}

// STRIDE threads share the groups of one decode region (the CTA's tile, or one warp's private part of it);
// idx is the thread's rank among them, t = e^{i 4 idx ratio} (times the region's offset into the tile),
// rstep = e^{i 4 STRIDE ratio}.
struct LeanParams {
    float2 g0;      // phasor of the thread's first sample (exact product)
    uint64_t m64k;  // see LeanPhase
    uint32_t mk32;
    float esc;
};

template <class Gm, int STRIDE, bool MIX>
__device__ __forceinline__ void decode_lean(const FirArgs &a, uint32_t raw_addr, uint32_t l_lo, uint32_t n_dec, uint64_t n0,
                                            const LeanParams &lp, float2 rstep, float4 *__restrict__ X4, int idx)
{
    static_assert(STRIDE % Gm::G == 0, "a thread's groups stay in one row");
    // local group gc holds region samples 4gc..4gc+3; a partial last group is decoded whole (its bytes are
    // inside the 16-byte-rounded copy and its slots inside the layout's slack; nothing reads them)
    const uint32_t n_loc = (n_dec + 3) >> 2;
    uint32_t rp = raw_addr + 8u * static_cast<uint32_t>(idx); // shared-window address of the thread's group
    const uint32_t rp_end = raw_addr + 8u * n_loc;
    float4 *xb = X4 + (idx & (Gm::G - 1)) * Gm::PITCH + (idx >> Gm::LOG_G);
    float2 g = make_float2(1.0f, 0.0f);
    uint64_t W = 0, wstep = 0;
    uint32_t mk32 = 0;
    float esc = 0.0f;
    if (MIX) {
        g = lp.g0;
        W = (n0 + static_cast<uint64_t>(4 * idx)) * lp.m64k;
        wstep = static_cast<uint64_t>(4 * STRIDE) * lp.m64k;
        mk32 = lp.mk32;
        esc = lp.esc;
    }
    const float2 negk = make_float2(-8388736.0f, -8388736.0f); // -(2^23 + 128)
    const float2 r1c = make_float2(a.rot[1].x, a.rot[1].x), r1s = make_float2(a.rot[1].y, a.rot[1].y);
    const float2 k2c = make_float2(a.rot2c, a.rot2c); // 2 cos(ratio)
    const float2 rsc = make_float2(rstep.x, rstep.x), rss = make_float2(rstep.y, rstep.y);
    if (l_lo) {
        // first tile of a launch that starts inside it: the groups below the first wanted sample are not there,
        // but the phasor and the rounding-error fraction step through them exactly as the full tile's loop would
        const uint32_t rp_first = raw_addr + 8u * (l_lo >> 2);
        for (; rp < rp_first && rp < rp_end; rp += 8u * STRIDE, xb += STRIDE / Gm::G) {
            if (MIX) {
                const float2 gp = make_float2(-g.y, g.x);
                g = fma2(gp, rss, mul2(g, rsc));
                W += wstep;
            }
        }
    }
#pragma unroll 4
    for (; rp < rp_end; rp += 8u * STRIDE, xb += STRIDE / Gm::G) {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(rp));
        const uint32_t u0 = v.x ^ 0x80808080u, u1 = v.y ^ 0x80808080u;
        float2 x[4];
        x[0] = add2(make_float2(__uint_as_float(__byte_perm(u0, 0x4B000000u, 0x7440)), __uint_as_float(__byte_perm(u0, 0x4B000000u, 0x7441))), negk);
        x[1] = add2(make_float2(__uint_as_float(__byte_perm(u0, 0x4B000000u, 0x7442)), __uint_as_float(__byte_perm(u0, 0x4B000000u, 0x7443))), negk);
        x[2] = add2(make_float2(__uint_as_float(__byte_perm(u1, 0x4B000000u, 0x7440)), __uint_as_float(__byte_perm(u1, 0x4B000000u, 0x7441))), negk);
        x[3] = add2(make_float2(__uint_as_float(__byte_perm(u1, 0x4B000000u, 0x7442)), __uint_as_float(__byte_perm(u1, 0x4B000000u, 0x7443))), negk);
        if (MIX) {
            const float2 gp = make_float2(-g.y, g.x); // i * g
            float2 ph[4];
            ph[0] = g;
            ph[1] = fma2(gp, r1s, mul2(g, r1c));
            // e^{i(n+1)w} = 2 cos w e^{inw} - e^{i(n-1)w}: one packed FMA each, two steps from exact anchors
            ph[2] = fma2(ph[1], k2c, make_float2(-g.x, -g.y));
            ph[3] = fma2(ph[2], k2c, make_float2(-ph[1].x, -ph[1].y));
            float e[4];
            const uint32_t w0 = static_cast<uint32_t>(W >> 32);
#pragma unroll
            for (int i = 0; i < 4; i += 2) { // signed fraction of an ulp -> radians, two samples per packed multiply
                const float2 ee = mul2(make_float2(static_cast<float>(static_cast<int>(w0 + static_cast<uint32_t>(i) * mk32)),
                                                   static_cast<float>(static_cast<int>(w0 + static_cast<uint32_t>(i + 1) * mk32))),
                                       make_float2(esc, esc));
                e[i] = ee.x, e[i + 1] = ee.y;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                // p = ph * (1 - i e), then (xr + i xi) * p = xr * p + xi * (i p): packed, with broadcast scalars
                const float2 p = fma2(make_float2(ph[i].y, -ph[i].x), make_float2(e[i], e[i]), ph[i]);
                x[i] = fma2(make_float2(-p.y, p.x), make_float2(x[i].y, x[i].y), mul2(p, make_float2(x[i].x, x[i].x)));
            }
            g = fma2(gp, rss, mul2(g, rsc));
            W += wstep;
        }
        xb[0] = make_float4(x[0].x, x[0].y, x[1].x, x[1].y);
        xb[Gm::G * Gm::PITCH] = make_float4(x[2].x, x[2].y, x[3].x, x[3].y);
    }
}

// ---------------------------------------------------------------------------- lean EXACT decode (integer formats)
// The same loop structure for the bit-exact mode: one 4-sample group per thread and iteration, incremental
// shared-memory addresses, integers made into floats by the mantissa trick (exact, like the conversion it
// replaces), then exactly the reference's operations: the IEEE quotient (div_exact2), the offset subtraction,
// fl64(n * ratio), an f64 sin/cos per sample and shift, and num-complex's multiply with every product and sum
// rounded on its own (packed: (x c, x s) + (-y s, y c) through an FMA by the opaque 1.0).
// a group of 4 raw samples held in words (cs16: 4 words, cs8/cu8: 2) -> the reference's decoded values
template <int FMT>
__device__ __forceinline__ void unpack_words(const uint32_t (&raw)[4], float2 (&x)[4], float2 one)
{
    if (FMT == QD_FMT_CS16) {
        const uint32_t w[4] = {raw[0] ^ 0x80008000u, raw[1] ^ 0x80008000u, raw[2] ^ 0x80008000u, raw[3] ^ 0x80008000u};
        const float2 negk = make_float2(-8421376.0f, -8421376.0f); // -(2^23 + 32768)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 n = add2(make_float2(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7410)),
                                              __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7432))), negk);
            x[i] = dec_fused2(n, 1.0f / 65535.0f, -32767.5f); // lib.rs:253
        }
    } else {
        const uint32_t flip = FMT == QD_FMT_CS8 ? 0x80808080u : 0u;
        const uint32_t w[2] = {raw[0] ^ flip, raw[1] ^ flip};
        const float kk = FMT == QD_FMT_CS8 ? -8388736.0f : -8388608.0f; // -(2^23 + 128) or -2^23
        const float2 negk = make_float2(kk, kk);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t word = w[i >> 1];
            const uint32_t sel = (i & 1) ? 0x7442 : 0x7440;
            const float2 n = add2(make_float2(__uint_as_float(__byte_perm(word, 0x4B000000u, sel)),
                                              __uint_as_float(__byte_perm(word, 0x4B000000u, sel + 1))), negk);
            if (FMT == QD_FMT_CS8) x[i] = div_exact2(n, 127.0f, 1.0f / 127.0f); // lib.rs:251
            else x[i] = dec_fused2(n, 1.0f / 255.0f, -127.5f); // lib.rs:252
        }
    }
}

template <int FMT>
__device__ __forceinline__ void unpack_group(uint32_t rp, float2 (&x)[4], float2 one)
{
    uint32_t w[4] = {0, 0, 0, 0};
    if (FMT == QD_FMT_CS16) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(rp));
    else asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(rp));
    unpack_words<FMT>(w, x, one);
}

// the reference's mixer on a group of four decoded samples: fl64(n * ratio), an f64 sin/cos per sample and shift
// (shift.rs:49-50), num-complex's multiply with every product and sum rounded on its own (shift.rs:51)
template <bool HALF = false>
__device__ __forceinline__ void mix_group_exact(const FirArgs &a, float2 (&x)[4], double nd)
{
    const float2 one = a.one;
    auto mix = [&](float2 v, double ni, double ratio) {
        double c, sn;
        sincos_f64k<HALF>(__dmul_rn(ni, ratio), a.sincos, a.k, c, sn);
        const float cf = static_cast<float>(c), sf = static_cast<float>(sn);
        const float2 p1 = mul2(make_float2(v.x, v.x), make_float2(cf, sf));
        const float2 p2 = mul2(make_float2(v.y, v.y), make_float2(-sf, cf));
        return fma2(p2, one, p1); // (x c - y s, x s + y c)
    };
    for (int sft = 0; sft < a.n_shift; sft++) {
        const double rs = a.ratio[sft];
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = mix(x[i], __dadd_rn(nd, static_cast<double>(i)), rs); // four independent chains
    }
}

// The lean EXACT decode for a tile that is NOT staged in shared memory (long filters: the staging buffer would cost
// a resident CTA): groups come straight from global memory -- an L2 hit after the bulk prefetch of the tile -- with
// the next group's load in flight while the current one is decoded and mixed.
template <class Gm, int STRIDE, int FMT>
__device__ __forceinline__ void decode_exact_global(const FirArgs &a, const uint8_t *__restrict__ g0, uint32_t n_dec, uint64_t n0,
                                                    float4 *__restrict__ X4, int idx, uint32_t g_first = 0)
{
    static_assert(STRIDE % Gm::G == 0, "a thread's groups stay in one row");
    const bool first_thread = idx == 0;
    idx += static_cast<int>(g_first); // groups below g_first are already in place (carried over from the CTA's previous tile)
    constexpr uint32_t GB = FMT == QD_FMT_CF32 ? 32u : (FMT == QD_FMT_CS16 ? 16u : 8u); // bytes per group of 4 samples
    // integer formats: a partial last group is decoded whole (its bytes lie inside the 16-byte granule the source
    // contract makes readable); cf32 groups are two granules, so the last 0..3 samples go one at a time
    const uint32_t n_loc = FMT == QD_FMT_CF32 ? (n_dec >> 2) : ((n_dec + 3) >> 2);
    float4 *xb = X4 + (idx & (Gm::G - 1)) * Gm::PITCH + (idx >> Gm::LOG_G);
    double nd = __ull2double_rn(n0 + static_cast<uint64_t>(4 * idx));
    constexpr int NW = FMT == QD_FMT_CF32 ? 8 : 4;
    auto fetch = [&](uint32_t gc, uint32_t (&w)[NW]) {
        if (FMT == QD_FMT_CF32) {
            const uint4 lo = ldg_stream_v4(g0 + static_cast<size_t>(GB) * gc), hi = ldg_stream_v4(g0 + static_cast<size_t>(GB) * gc + 16);
            w[0] = lo.x, w[1] = lo.y, w[2] = lo.z, w[3] = lo.w;
            w[NW - 4] = hi.x, w[NW - 3] = hi.y, w[NW - 2] = hi.z, w[NW - 1] = hi.w;
        } else if (FMT == QD_FMT_CS16) {
            const uint4 v = ldg_stream_v4(g0 + static_cast<size_t>(GB) * gc);
            w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
        } else {
            const uint2 v = ldg_stream_v2(g0 + static_cast<size_t>(GB) * gc);
            w[0] = v.x, w[1] = v.y, w[2] = 0, w[3] = 0;
        }
    };
    uint32_t nxt[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) nxt[i] = 0;
    if (static_cast<uint32_t>(idx) < n_loc) fetch(idx, nxt);
#pragma unroll 2
    for (uint32_t gc = idx; gc < n_loc; gc += STRIDE, xb += STRIDE / Gm::G) {
        uint32_t cur[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) cur[i] = nxt[i];
        if (gc + STRIDE < n_loc) fetch(gc + STRIDE, nxt);
        float2 x[4];
        if constexpr (FMT == QD_FMT_CF32) {
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = make_float2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])); // bit copy, lib.rs:248
        } else {
            const uint32_t c4[4] = {cur[0], cur[1], cur[2], cur[3]};
            unpack_words<FMT>(c4, x, a.one);
        }
        if (a.n_shift) {
            mix_group_exact<(FMT == QD_FMT_CF32)>(a, x, nd); // (integer tiles of the long-filter kernels: the full table)
            nd = __dadd_rn(nd, static_cast<double>(4 * STRIDE));
        }
        xb[0] = make_float4(x[0].x, x[0].y, x[1].x, x[1].y);
        xb[Gm::G * Gm::PITCH] = make_float4(x[2].x, x[2].y, x[3].x, x[3].y);
    }
    if (FMT == QD_FMT_CF32 && first_thread) { // the last 0..3 samples, one at a time
        for (uint32_t l = 4 * n_loc; l < n_dec; l++) {
            float2 v = __ldg(reinterpret_cast<const float2 *>(g0) + l);
            const double ni = __ull2double_rn(n0 + l);
            for (int sft = 0; sft < a.n_shift; sft++) v = mix_exact<false>(v, ni, a.ratio[sft], a);
            const uint32_t pr = (l >> 1) & (Gm::DR / 2 - 1);
            float4 *el = X4 + ((pr & 1) * Gm::G + (pr >> 1)) * Gm::PITCH + (l >> Gm::LOG_DR);
            reinterpret_cast<float2 *>(el)[l & 1] = v;
        }
    }
}

template <class Gm, int STRIDE, int FMT>
__device__ __forceinline__ void decode_lean_exact(const FirArgs &a, uint32_t raw_addr, uint32_t n_dec, uint64_t n0,
                                                  float4 *__restrict__ X4, int idx)
{
    static_assert(STRIDE % Gm::G == 0, "a thread's groups stay in one row");
    constexpr uint32_t GB = FMT == QD_FMT_CS16 ? 16u : 8u; // bytes per group of 4 samples
    const uint32_t n_loc = (n_dec + 3) >> 2;
    uint32_t rp = raw_addr + GB * static_cast<uint32_t>(idx);
    const uint32_t rp_end = raw_addr + GB * n_loc;
    float4 *xb = X4 + (idx & (Gm::G - 1)) * Gm::PITCH + (idx >> Gm::LOG_G);
    const float2 one = a.one;
    const int n_shift = a.n_shift;
    double nd = __ull2double_rn(n0 + static_cast<uint64_t>(4 * idx)); // absolute index of the group's first sample
#pragma unroll 2
    for (; rp < rp_end; rp += GB * STRIDE, xb += STRIDE / Gm::G) {
        float2 x[4];
        unpack_group<FMT>(rp, x, one);
        if (n_shift) {
            auto mix = [&](float2 v, double ni, double ratio) {
                double c, sn;
                sincos_f64k(__dmul_rn(ni, ratio), a.sincos, a.k, c, sn); // shift.rs:49-50
                const float cf = static_cast<float>(c), sf = static_cast<float>(sn);
                const float2 p1 = mul2(make_float2(v.x, v.x), make_float2(cf, sf));
                const float2 p2 = mul2(make_float2(v.y, v.y), make_float2(-sf, cf));
                return fma2(p2, one, p1); // (x c - y s, x s + y c), shift.rs:51
            };
            const double r0 = a.ratio[0];
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = mix(x[i], __dadd_rn(nd, static_cast<double>(i)), r0); // four independent chains
            for (int sft = 1; sft < n_shift; sft++) {
                const double rs = a.ratio[sft];
#pragma unroll
                for (int i = 0; i < 4; i++) x[i] = mix(x[i], __dadd_rn(nd, static_cast<double>(i)), rs);
            }
            nd = __dadd_rn(nd, static_cast<double>(4 * STRIDE));
        }
        xb[0] = make_float4(x[0].x, x[0].y, x[1].x, x[1].y);
        xb[Gm::G * Gm::PITCH] = make_float4(x[2].x, x[2].y, x[3].x, x[3].y);
    }
}

// ---------------------------------------------------------------------------- cf32 without a shift: a plain copy
// A cf32 sample's "decode" is a bit copy (lib.rs:248), so with no shift in front of the filter a group of four
// samples goes from global memory to its two slots of the polyphase layout untouched.  With 8 bytes per sample
// this path is bound by the bytes it has in flight: two groups (four 128-bit loads) per thread are issued
// before the first store.
template <class Gm, int STRIDE>
__device__ __forceinline__ void decode_cf32_copy(const uint8_t *gsrc, uint32_t n_dec, float2 *__restrict__ X, int idx)
{
    static_assert(STRIDE % Gm::G == 0, "a thread's groups stay in one row");
    constexpr int B = 2; // groups in flight per thread (three or four spill under the 96-register cap and measure slower)
    const uint32_t n_full = n_dec >> 2;
    const uint4 *gp = reinterpret_cast<const uint4 *>(gsrc) + 2u * static_cast<uint32_t>(idx);
    float4 *xb = reinterpret_cast<float4 *>(X) + (idx & (Gm::G - 1)) * Gm::PITCH + (idx >> Gm::LOG_G);
    for (uint32_t gc = idx; gc < n_full; gc += B * STRIDE, gp += 2u * B * STRIDE, xb += B * (STRIDE / Gm::G)) {
        uint4 v[2 * B];
#pragma unroll
        for (int j = 0; j < B; j++) {
            if (gc + j * STRIDE < n_full) {
                v[2 * j] = __ldg(gp + 2u * j * STRIDE);
                v[2 * j + 1] = __ldg(gp + 2u * j * STRIDE + 1);
            }
        }
#pragma unroll
        for (int j = 0; j < B; j++) {
            if (gc + j * STRIDE < n_full) {
                float4 *o = xb + j * (STRIDE / Gm::G);
                o[0] = make_float4(__uint_as_float(v[2 * j].x), __uint_as_float(v[2 * j].y), __uint_as_float(v[2 * j].z), __uint_as_float(v[2 * j].w));
                o[Gm::G * Gm::PITCH] = make_float4(__uint_as_float(v[2 * j + 1].x), __uint_as_float(v[2 * j + 1].y), __uint_as_float(v[2 * j + 1].z), __uint_as_float(v[2 * j + 1].w));
            }
        }
    }
    if (idx == 0) { // the last 0..3 samples, one at a time
        for (uint32_t l = 4 * n_full; l < n_dec; l++) {
            const uint32_t pr = (l >> 1) & (Gm::DR / 2 - 1);
            float4 *el = reinterpret_cast<float4 *>(X) + ((pr & 1) * Gm::G + (pr >> 1)) * Gm::PITCH + (l >> Gm::LOG_DR);
            reinterpret_cast<float2 *>(el)[l & 1] = __ldg(reinterpret_cast<const float2 *>(gsrc) + l);
        }
    }
}

template <int D, int R, int NT, int LMAX, bool MIX>
__device__ __forceinline__ void decode_tile_lean(const FirArgs &a, const uint8_t *raw, uint32_t lead, uint32_t l_lo, uint32_t n_dec,
                                                 uint64_t n_tile0, const LeanPhase *lp, const double2 *ttab,
                                                 float2 *__restrict__ X, int tid)
{
    LeanParams q;
    q.g0 = make_float2(1.0f, 0.0f);
    q.m64k = 0, q.mk32 = 0, q.esc = 0.0f;
    if (MIX) {
        const double2 t = ttab[tid];
        const double ac = lp->ac, as = lp->as;
        q.g0 = make_float2(static_cast<float>(fma(ac, t.x, -__dmul_rn(as, t.y))), static_cast<float>(fma(ac, t.y, __dmul_rn(as, t.x))));
        q.m64k = lp->m64k, q.mk32 = lp->mk32, q.esc = lp->esc;
    }
    decode_lean<FirGeom<D, R, NT, LMAX>, NT, MIX>(a, smem_u32(raw) + 8u * (lead >> 2), l_lo, n_dec, n_tile0, q, a.rot_step,
                                                         reinterpret_cast<float4 *>(X), tid);
}

// ---------------------------------------------------------------------------- FIR stage
// Block b of a thread = its local samples s = b*D .. b*D+D-1: pair rows (b mod R)*D/2 + p/2, column tid + b div R.
template <int D, int R, int NT, int LMAX>
__device__ __forceinline__ void load_block(const float2 *__restrict__ xcol, int rb, float2 (&v)[D])
{
    using Gm = FirGeom<D, R, NT, LMAX>;
    static_assert(D % 2 == 0, "pairs of samples");
    const float4 *xc4 = reinterpret_cast<const float4 *>(xcol);
#pragma unroll
    for (int p = 0; p < D; p += 2) {
        const int r = rb * (D / 2) + p / 2;
        const float4 q = xc4[((r & 1) * Gm::G + (r >> 1)) * Gm::PITCH];
        v[p] = make_float2(q.x, q.y);
        v[p + 1] = make_float2(q.z, q.w);
    }
}

// every check at run time: prologue / epilogue blocks, partial tap blocks, truncated reads
template <int D, int R, int NT, int LMAX, bool EXACT>
__device__ __forceinline__ void general_block(const float2 *__restrict__ X, int tid, int b, int s_end, int Q, int Lrem,
                                              const FirTaps &taps, float2 one, float2 (&acc)[R])
{
    const int plim = s_end - b * D;
    if (plim <= 0) return;
    float2 v[D];
    load_block<D, R, NT, LMAX>(X + 2 * (tid + b / R), b & (R - 1), v);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int qb = b - r; // tap block of output r at this step
        if (qb < 0 || qb >= Q) continue;
        const int pmax = min(plim, qb == Q - 1 ? Lrem : D);
        const float2 *tp = taps.t + qb * D;
        if (pmax >= D) {
#pragma unroll
            for (int p = 0; p < D; p++) acc[r] = mac<EXACT>(acc[r], v[p], tp[p], one);
        } else {
#pragma unroll
            for (int p = 0; p < D; p++)
                if (p < pmax) acc[r] = mac<EXACT>(acc[r], v[p], tp[p], one);
        }
    }
}

// LS > 0: the filter length is a compile-time constant and the whole tap schedule unrolls.
// TRUNC: the same schedule with every MAC predicated on its sample existing for this read (s < s_end): what a
// warp runs when one of its threads owns the truncated tail of a unit, instead of sending that one thread
// through the general loop while the other 31 wait.
// TAILS: output r's accumulator is copied to snap[r] when it holds exactly snap_j[r] taps (see FirArgs::tail_out);
// the tap counts a snapshot can ask for, L/2 + m*D, are known at compile time, so only those places carry a test.
template <int D, int R, int NT, int LMAX, bool EXACT, int LS, bool TRUNC, bool TAILS = false>
__device__ __forceinline__ void fir_static(const float2 *__restrict__ X, int tid, const FirTaps &taps, float2 one,
                                           float2 (&acc)[R], int s_end, const int *snap_j = nullptr, float2 *snap = nullptr)
{
    constexpr int Q = (LS + D - 1) / D, LREM = LS - (Q - 1) * D, NB = R - 1 + Q;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        float2 v[D];
        load_block<D, R, NT, LMAX>(X + 2 * (tid + b / R), b % R, v);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int qb = b - r;
            if (qb < 0 || qb >= Q) continue;
#pragma unroll
            for (int p = 0; p < D; p++)
                if (p < (qb == Q - 1 ? LREM : D) && (!TRUNC || b * D + p < s_end)) {
                    constexpr int half = LS / 2;
                    const int j = qb * D + p; // taps already applied to output r
                    if (TAILS && j > half && (j - half) % D == 0 && snap_j[r] == j) snap[r] = acc[r];
                    acc[r] = mac<EXACT>(acc[r], v[p], taps.t[qb * D + p], one);
                }
        }
    }
}

// Run-time filter length.  Unit (b, r) = sample block b of the thread applied to its output r with tap block
// q = b - r; it exists for 0 <= q < Q and is whole for q < Qf = L / D.  Every block number, tap index and trip count
// below is a function of kernel parameters alone, so the taps come from the constant bank through UNIFORM
// registers (LDCU, four taps per load, broadcast into the packed multiplies: no per-thread constant loads) -- the
// CPU-side test tests/test_sass.py checks that ptxas keeps it that way.  Blocks R-1 <= b < Qf carry a whole tap
// block for every output: one block per iteration, the next block's samples in flight while this one is applied.
// The few blocks before and after run the same way with the units that exist.  A thread whose unit ends early
// (the zero-truncated tail of a read, filter.rs:68-71) stays in the loops with its accumulator updates predicated
// off block by block; if its last block is a partial one it is applied afterwards (nothing follows it in that
// thread's sums).
// snap_blk (SNAP): output r's accumulator is copied to snap[r] right before block snap_blk[r] of the thread is
// applied (see FirArgs::tail_out); the caller completes a snapshot that ends inside that block.
template <int D, int R, int NT, int LMAX, bool EXACT, bool SNAP = false>
__device__ __forceinline__ void fir_dynamic(const float2 *__restrict__ X, int tid, int Q, int Lrem, int s_end,
                                            const FirTaps &taps, float2 one, float2 (&acc)[R], const int *snap_blk = nullptr,
                                            float2 *snap = nullptr)
{
    auto take = [&](int blk) { // only stream launches that feed overlapping windows carry snapshots
        if constexpr (SNAP) {
#pragma unroll
            for (int r = 0; r < R; r++)
                if (snap_blk[r] == blk) snap[r] = acc[r];
        }
    };
    // s_end: the thread's samples s >= s_end do not exist for this read; untruncated threads pass (R-1)*D + L
    const int NB = R - 1 + Q;                // blocks of a thread
    const int Qf = Lrem == D ? Q : Q - 1;    // whole tap blocks
    const int part = Lrem == D ? 0 : Lrem;   // taps of the partial last tap block
    const int nfull = s_end / D;             // the thread's blocks b < nfull are whole
    // a whole tap block on one output
    auto whole = [&](float2 &t, const float2 (&v)[D], int qb) {
        const float *ts = taps.s + qb * D;
        if (D % 4 == 0) {
#pragma unroll
            for (int p = 0; p < D; p += 4) {
                const float4 f = *reinterpret_cast<const float4 *>(ts + p);
                t = mac<EXACT>(t, v[p], make_float2(f.x, f.x), one);
                t = mac<EXACT>(t, v[p + 1], make_float2(f.y, f.y), one);
                t = mac<EXACT>(t, v[p + 2], make_float2(f.z, f.z), one);
                t = mac<EXACT>(t, v[p + 3], make_float2(f.w, f.w), one);
            }
        } else {
#pragma unroll
            for (int p = 0; p < D; p++) t = mac<EXACT>(t, v[p], make_float2(ts[p], ts[p]), one);
        }
    };
    // a block in which not every output has a (whole) tap block: the first R - 1 and the last few
    auto edge_block = [&](int blk, int rb) {
        take(blk);
        float2 v[D];
        load_block<D, R, NT, LMAX>(X + 2 * (tid + blk / R), rb, v);
        const bool on = blk < nfull;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int qb = blk - r; // uniform
            if (qb < 0 || qb >= Q) continue;
            float2 t = acc[r];
            if (qb < Qf) {
                whole(t, v, qb);
            } else {
                const float *ts = taps.s + qb * D;
#pragma unroll
                for (int p = 0; p < D; p++)
                    if (p < part) t = mac<EXACT>(t, v[p], make_float2(ts[p], ts[p]), one);
            }
            if (on) acc[r] = t;
        }
    };
#pragma unroll
    for (int bb = 0; bb < R - 1; ++bb) edge_block(bb, bb); // (bb < NB: every filter has at least one tap block)
    int b = R - 1;
    if (b < Qf) {
        // the block's row group (b mod R) and column (b div R) are stepped, never divided
        const float4 *xc4 = reinterpret_cast<const float4 *>(X + 2 * tid);
        int rb = R - 1;
        // applies block blk held in v[] to all R outputs (whole tap blocks for every output)
        // the per-thread tests run on per-thread countdowns: comparing the block counter itself with a thread's
        // values would pull it, and with it the tap indices, out of the uniform registers
        int left = nfull - b; // whole blocks the thread still has
        int until[R];         // blocks until output r's snapshot
#pragma unroll
        for (int r = 0; r < R; r++) until[r] = SNAP ? snap_blk[r] - b : 0;
        auto apply = [&](const float2 (&v)[D], int blk) {
            const bool on = left > 0;
            left--;
            if constexpr (SNAP) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (until[r] == 0) snap[r] = acc[r];
                    until[r]--;
                }
            }
            float2 t[R]; // the block's sums are kept or dropped as a whole: R selects per block, none per MAC
#pragma unroll
            for (int r = 0; r < R; r++) t[r] = acc[r];
            if (D % 4 == 0) {
#pragma unroll
                for (int p = 0; p < D; p += 4) {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const float4 f = *reinterpret_cast<const float4 *>(taps.s + (blk - r) * D + p);
                        t[r] = mac<EXACT>(t[r], v[p], make_float2(f.x, f.x), one);
                        t[r] = mac<EXACT>(t[r], v[p + 1], make_float2(f.y, f.y), one);
                        t[r] = mac<EXACT>(t[r], v[p + 2], make_float2(f.z, f.z), one);
                        t[r] = mac<EXACT>(t[r], v[p + 3], make_float2(f.w, f.w), one);
                    }
                }
            } else {
#pragma unroll
                for (int p = 0; p < D; p++) {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const float f = taps.s[(blk - r) * D + p];
                        t[r] = mac<EXACT>(t[r], v[p], make_float2(f, f), one);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++)
                if (on) acc[r] = t[r];
        };
        auto step = [&]() { // to the next block's place in the layout
            if (++rb == R) {
                rb = 0;
                xc4 += 1;
            }
        };
        // two blocks per iteration in two register sets: each set is loaded a whole block ahead of its use and no
        // copy between the sets ever waits for a load (the block after the last is inside the layout: Qf <= NB - 1)
        float2 v0[D], v1[D];
        load_block<D, R, NT, LMAX>(reinterpret_cast<const float2 *>(xc4), rb, v0);
        // (an explicit trip count: with `b + 2 <= Qf` as the exit test ptxas moves b to a vector register)
        const int n2 = (Qf - (R - 1)) >> 1;
        for (int it = 0; it < n2; ++it) {
            const int blk = R - 1 + 2 * it; // from the trip counter alone: b itself has per-thread uses after the loop
            step();
            load_block<D, R, NT, LMAX>(reinterpret_cast<const float2 *>(xc4), rb, v1);
            apply(v0, blk);
            step();
            load_block<D, R, NT, LMAX>(reinterpret_cast<const float2 *>(xc4), rb, v0);
            apply(v1, blk + 1);
        }
        b = R - 1 + 2 * n2;
        if ((Qf - (R - 1)) & 1) {
            apply(v0, b);
            ++b;
        }
    }
    // the last blocks: b is R - 1 or Qf here, whichever is larger
#pragma unroll
    for (int m = 0; m < R; ++m)
        if (b + m < NB) edge_block(b + m, (b + m) % R);
    // a truncated thread's partial last block (per-thread path: its tap count is the thread's own)
    if (s_end - nfull * D > 0 && nfull < NB) general_block<D, R, NT, LMAX, EXACT>(X, tid, nfull, s_end, Q, Lrem, taps, one, acc);
}

// FIR + store of one tile: thread `tid` of the tile owns outputs R*tid .. R*tid+R-1; its samples sit in the
// layout X (geometry FirGeom<D, R, NTG, LMAX>) from column xidx on
// FUSE != 0: the outputs are not stored; they come back in acc[] (valid where `mine`) for the fused STFT, and with
// FUSE == 2 the snapshots in snapv[] (bit r of snap_mask: output r has one)
template <int D, int R, int NTG, int LMAX, bool EXACT, int LS, bool SNAP, int FUSE = 0>
__device__ __forceinline__ bool fir_tile(const FirArgs &a, const FirTaps &taps, const TileGeo &g, const float2 *__restrict__ X,
                                 int xidx, int tid, float2 (&acc)[R], float2 (&snapv)[R], uint32_t &snap_mask)
{
    snap_mask = 0;
    // slots [skip, cnt) of the tile are wanted; a thread takes part when any of its R slots is.  Whole WARPS enter
    // (a vote, not a per-thread branch: the filter loops then run in warp-uniform control flow, which is what lets
    // ptxas keep their counters and tap indices in uniform registers); a thread without slots runs them predicated
    // off and stores nothing.
    const bool mine = static_cast<uint32_t>(R * tid) < g.cnt && static_cast<uint32_t>(R * tid + R) > g.skip;
    if (__any_sync(0xffffffffu, mine)) {
        const int64_t q = g.f0 + static_cast<int64_t>(R * tid); // the thread's first slot as a flat output index
        // the unit's raw buffer ends at (unit_top0 + n_call)*D + L: later samples do not exist for
        // this read (filter.rs:68-71) and the ascending tap loop stops there
        int64_t s_lim; // samples from the thread's first one to the end of its unit's raw buffer
        if (a.contiguous && a.ncall_log2 >= 0) {
            // units tile the output stream and n_call is a power of two: the outputs left in the thread's unit
            // come from a mask, and raw_end - n_first = left * D + (L - i0)
            const uint64_t qq = q < 0 ? 0 : static_cast<uint64_t>(q);
            const uint64_t left = a.n_call - (qq & (a.n_call - 1)) + (qq - static_cast<uint64_t>(q));
            s_lim = static_cast<int64_t>(left * D + (a.L - (a.L - a.L / 2)));
        } else {
            uint64_t unit_top0;
            if (a.contiguous) {
                // one division per tile (g.u0 = f0 / n_call); a thread's outputs start `rel` past that unit
                int64_t rel = q - static_cast<int64_t>(g.u0 * a.n_call);
                uint64_t un = g.u0;
                while (rel >= static_cast<int64_t>(a.n_call)) { // the tile runs over one or more unit boundaries
                    rel -= static_cast<int64_t>(a.n_call);
                    un++;
                }
                unit_top0 = a.off0 + un * a.n_call;
            } else {
                unit_top0 = a.off0 + g.unit * a.S;
            }
            const uint64_t raw_end = (unit_top0 + a.n_call) * D + a.L;
            s_lim = static_cast<int64_t>(raw_end - g.n_tile0) - static_cast<int64_t>(tid) * (D * R);
        }
        const int L = LS > 0 ? LS : static_cast<int>(a.L);
        const int s_total = (R - 1) * D + L;
        const int Q = (L + D - 1) / D, Lrem = L - (Q - 1) * D;
        const float2 one = a.one;

    #pragma unroll
        for (int r = 0; r < R; r++) acc[r] = make_float2(0.0f, 0.0f); // Complex::zero(), filter.rs:112

        // the tail of a read: outputs whose taps run past the end of the unit's raw buffer stop there
        const int s_end = mine ? static_cast<int>(min(static_cast<int64_t>(s_total), s_lim)) : 0;
        if constexpr (!SNAP) {
            if (LS > 0) {
                if (!__any_sync(__activemask(), s_lim < s_total)) fir_static<D, R, NTG, LMAX, EXACT, (LS > 0 ? LS : 1), false>(X, xidx, taps, one, acc, s_end);
                else fir_static<D, R, NTG, LMAX, EXACT, (LS > 0 ? LS : 1), true>(X, xidx, taps, one, acc, s_end);
            } else {
                fir_dynamic<D, R, NTG, LMAX, EXACT>(X, xidx, Q, Lrem, s_end, taps, one, acc);
            }
        } else {
            // stream feeding overlapping windows: some outputs also leave a snapshot of their running sum
            int snap_j[R], snap_blk[R];
            uint64_t snap_at[R];
            float2 snap[R];
            const int64_t lead_in = static_cast<int64_t>(a.tail_W - a.tail_T);
#pragma unroll
            for (int r = 0; r < R; r++) {
                snap_j[r] = snap_blk[r] = -1;
                snap_at[r] = 0;
                snap[r] = make_float2(0.0f, 0.0f);
                const int64_t qq = q + r - lead_in;
                if (a.tail_T != 0 && mine && qq >= 0 && q + r < static_cast<int64_t>(a.total_out)) {
                    const uint64_t u = static_cast<uint64_t>(qq) / a.tail_S;
                    const uint32_t rr = static_cast<uint32_t>(static_cast<uint64_t>(qq) - u * a.tail_S);
                    if (rr < a.tail_T && u < a.tail_units) {
                        snap_j[r] = static_cast<int>((a.tail_T - rr) * D) + L / 2; // taps of that window's output
                        snap_blk[r] = snap_j[r] / D + r;                            // the thread's block they end in
                        snap_at[r] = u * a.tail_T + rr;
                    }
                }
            }
            if (LS > 0) fir_static<D, R, NTG, LMAX, EXACT, (LS > 0 ? LS : 1), false, true>(X, xidx, taps, one, acc, s_end, snap_j, snap);
            else fir_dynamic<D, R, NTG, LMAX, EXACT, true>(X, xidx, Q, Lrem, s_end, taps, one, acc, snap_blk, snap);
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (snap_j[r] < 0) continue;
                if (LS == 0) { // the taps of the snapshot's last, partial block
                    const int bs = snap_blk[r], p0 = snap_j[r] - (bs - r) * D;
                    if (p0 > 0) {
                        float2 v[D];
                        load_block<D, R, NTG, LMAX>(X + 2 * (xidx + bs / R), bs % R, v);
                        const float *tp = taps.s + (bs - r) * D;
#pragma unroll
                        for (int p = 0; p < D; p++)
                            if (p < p0) snap[r] = mac<EXACT>(snap[r], v[p], make_float2(tp[p], tp[p]), one);
                    }
                }
                if (FUSE == 2) {
                    snapv[r] = snap[r];
                    snap_mask |= 1u << r;
                } else {
                    a.tail_out[snap_at[r]] = snap[r];
                }
            }
        }
        float2 *o = a.out + q;
        if (FUSE || !mine) {
            // no slot of this thread is wanted, or the caller takes the outputs from acc[]
        } else if (q >= 0 && q + R <= static_cast<int64_t>(a.total_out)) {
            if (R % 2 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                for (int r = 0; r < R; r += 2)
                    *reinterpret_cast<float4 *>(o + r) = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) o[r] = acc[r];
            }
        } else { // a launch that starts or ends inside the thread's slots
#pragma unroll
            for (int r = 0; r < R; r++)
                if (q + r >= 0 && q + r < static_cast<int64_t>(a.total_out)) o[r] = acc[r];
        }
    }
    return mine;
}

// ---------------------------------------------------------------------------- fused sparkfft sink
// The tile's outputs (R per thread, in registers) are whole windows of W = a.fft.W samples.  They are laid into
// shared memory (the sample layout X is dead once every thread has left the filter) in the leaf order of our
// radix-4 FFT, transformed in place -- the passes, butterflies and twiddles of gk_fft, hence of the oracle, bit for
// bit -- and leave as glyph indices / magnitudes (fft.rs:48-60).  The STFT is a few per cent of the filter's work,
// so its passes are plain cooperative loops over the tile.
template <int R, int NT, int T_TILE>
__device__ __forceinline__ void fused_stft(const FirArgs &a, const TileGeo &g, float2 *__restrict__ Y, const float2 (&acc)[R],
                                           bool mine, int tid)
{
    const uint32_t W = a.fft.W;
    const int logw = 31 - __clz(W);
    const bool odd = logw & 1;
    const int n_r4 = logw >> 1;
    if (mine) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t slot = static_cast<uint32_t>(R * tid + r);
            if (slot >= g.skip && slot < g.cnt) Y[(slot & ~(W - 1)) + leaf_position(slot & (W - 1), W, n_r4, odd)] = acc[r];
        }
    }
    __syncthreads();
    if (odd) { // innermost size-2 FFTs (windows are W-aligned runs of the tile, so pairs never straddle one)
        for (uint32_t b = tid; b < T_TILE / 2; b += NT) {
            const float2 p = Y[2 * b], q = Y[2 * b + 1];
            Y[2 * b] = padd(p, q);
            Y[2 * b + 1] = psub(p, q);
        }
        __syncthreads();
    }
    int logq = odd ? 1 : 0;
    for (uint32_t q = odd ? 2 : 1; q < W; q <<= 2, logq += 2) {
        const uint32_t scale = W >> (logq + 2); // w(4q, j) == w(W, j * W/(4q))
        for (uint32_t b = tid; b < T_TILE / 4; b += NT) {
            const uint32_t wdw = b >> (logw - 2), bb = b & ((W >> 2) - 1);
            const uint32_t blk = bb >> logq, k = bb & (q - 1);
            float2 *base = Y + (wdw << logw) + (blk << (logq + 2)) + k;
            float2 t0 = base[0], t1 = base[q], t2 = base[2 * q], t3 = base[3 * q];
            if (k != 0) { // packed (re, im) arithmetic: every half is the individually rounded scalar operation
                t1 = pmul_tw(t1, __ldg(a.fft.tw + k * scale), a.one);
                t2 = pmul_tw(t2, __ldg(a.fft.tw + 2 * k * scale), a.one);
                t3 = pmul_tw(t3, __ldg(a.fft.tw + 3 * k * scale), a.one);
            }
            pradix4(t0, t1, t2, t3);
            base[0] = t0;
            base[q] = t1;
            base[2 * q] = t2;
            base[3 * q] = t3;
        }
        __syncthreads();
    }
    // epilogue: thread t takes bins R*t .. R*t+R-1 (FFT order) of its window; their display places (bins W/2..W-1,
    // then 0..W/2-1) are R consecutive bytes of the row when the window is at least 2R wide, so the glyphs leave as
    // one word (index-only output through the threshold test); anything else goes bin by bin
    static_assert(T_TILE == R * NT, "one run of R slots per thread");
    const uint32_t i0 = static_cast<uint32_t>(R * tid);
    if (i0 < g.skip || i0 >= g.cnt) return;
    const uint64_t u = static_cast<uint64_t>(g.f0 + static_cast<int64_t>(i0 & ~(W - 1))) >> logw; // the window's row
    const uint32_t pos0 = i0 & (W - 1);
    uint8_t *row = a.fft.idx + u * W + ((pos0 + W / 2) & (W - 1));
    if ((R == 4 || R == 8) && W >= 2 * R && !a.fft.mag && a.fft.use_thr && (reinterpret_cast<uintptr_t>(row) & 3) == 0) {
#pragma unroll
        for (int h = 0; h < R; h += 4)
            *reinterpret_cast<uint32_t *>(row + h) = glyph4_word(a.fft, Y[i0 + h], Y[i0 + h + 1], Y[i0 + h + 2], Y[i0 + h + 3]);
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t i = i0 + r;
            if (i >= g.cnt) break;
            const uint64_t ui = static_cast<uint64_t>(g.f0 + static_cast<int64_t>(i & ~(W - 1))) >> logw;
            emit_bin(a.fft, ui, W, i & (W - 1), Y[i]);
        }
    }
}

// The same for OVERLAPPING windows cut from a stream (stride S < width W; FUSE = 2).  A window is transformed by
// the tile that holds its last output; the W - 1 outputs before the tile that it may also need come from the
// carry buffers, which the CTA fills at the end of every tile (a CTA walks a contiguous run of tiles, and first
// filters the tile before its run just to have them: `warm`).  The last T samples of a window are the snapshots of
// the same stream outputs (FirArgs::tail_out), carried the same way.  Scratch inside the dead sample layout:
// values and snapshots of carry + tile, then one W-point work array per window of the tile.
template <int R, int NT, int T_TILE>
__device__ __forceinline__ void fused_stream_stft(const FirArgs &a, const TileGeo &g, float2 *__restrict__ scratch,
                                                  float2 *__restrict__ carry, const float2 (&acc)[R], const float2 (&snapv)[R],
                                                  uint32_t snap_mask, bool mine, bool warm, int tid)
{
    const uint32_t W = a.fft.W, S = a.fuse_S, T = a.tail_T, L2 = a.st2_L, D2 = a.st2_D;
    const uint32_t C1 = L2 ? L2 - 1 : W - 1; // carried outputs of this launch's stream
    const uint32_t C2 = L2 ? W - 1 : 0;      // carried second-stage outputs
    const uint32_t N2 = L2 ? T_TILE / D2 + 2 : 0;
    const int logw = 31 - __clz(W);
    const bool odd = logw & 1;
    const int n_r4 = logw >> 1;
    // (stream outputs sit at ysk(index): one slot of skew per 32, so that the second stage's threads, D2 outputs
    // apart, do not all read one bank)
    auto ysk = [](uint32_t y) { return y + (y >> 5); };
    const uint32_t YN = ysk(C1 + T_TILE) + 1;
    float2 *Yv = scratch, *Ys = Yv + YN, *Zv = Ys + YN, *Wk = Zv + (C2 + N2);
    float2 *Cv = carry, *Cs = Cv + C1, *Cz = Cs + C1;
    // 1. carry + this tile's outputs and snapshots, indexed by (flat output index - f0 + C1)
    for (uint32_t i = tid; i < C1; i += NT) {
        Yv[ysk(i)] = Cv[i];
        Ys[ysk(i)] = Cs[i];
    }
    for (uint32_t i = tid; i < C2; i += NT) Zv[i] = Cz[i];
    if (mine) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t slot = static_cast<uint32_t>(R * tid + r);
            if (slot >= g.skip && slot < g.cnt) {
                Yv[ysk(C1 + slot)] = acc[r];
                if (snap_mask & (1u << r)) Ys[ysk(C1 + slot)] = snapv[r];
            }
        }
    }
    __syncthreads();
    // (32-bit index arithmetic: the host fuses only launches of fewer than 2^31 stream outputs)
    const int f0 = static_cast<int>(g.f0);
    const int first_flat = f0 + static_cast<int>(g.skip); // stream outputs this tile adds
    const int last_flat = f0 + static_cast<int>(g.cnt) - 1;
    // what the windows are cut from: this stream (with snapshots for their truncated tails) or the second stage
    const bool from_y = L2 == 0;
    int src_base = f0 - static_cast<int>(C1);   // flat index of element 0 of the source
    int new_lo = first_flat, new_hi = last_flat; // flat range of the source elements this tile adds
    uint32_t n2 = 0;
    if (L2) {
        // second-stage outputs whose last input is one of this tile's outputs: k*D2 + L2 - 1 in [first_flat, last_flat]
        int k_lo = first_flat - static_cast<int>(L2 - 1);
        k_lo = k_lo <= 0 ? 0 : (k_lo + static_cast<int>(D2) - 1) / static_cast<int>(D2);
        int k_hi = last_flat - static_cast<int>(L2 - 1);
        k_hi = k_hi < 0 ? -1 : k_hi / static_cast<int>(D2);
        if (k_hi >= static_cast<int>(a.st2_total)) k_hi = static_cast<int>(a.st2_total) - 1;
        n2 = k_hi >= k_lo ? static_cast<uint32_t>(k_hi - k_lo + 1) : 0u;
        for (uint32_t i = tid; i < n2; i += NT) {
            const uint32_t y0 = static_cast<uint32_t>((k_lo + static_cast<int>(i)) * static_cast<int>(D2) - f0 + static_cast<int>(C1));
            float2 z = make_float2(0.0f, 0.0f); // ascending taps, product and sum rounded separately (filter.rs:112-120)
            uint32_t j = 0;
            for (; j + 8 <= L2; j += 8) { // eight loads in flight; the sum itself stays one chain in tap order
                float2 v[8];
                float t[8];
#pragma unroll
                for (int e = 0; e < 8; e++) v[e] = Yv[ysk(y0 + j + e)], t[e] = a.st2_taps[j + e];
#pragma unroll
                for (int e = 0; e < 8; e++) z = fma2(mul2(v[e], make_float2(t[e], t[e])), a.one, z);
            }
            for (; j < L2; j++) z = fma2(mul2(Yv[ysk(y0 + j)], make_float2(a.st2_taps[j], a.st2_taps[j])), a.one, z);
            Zv[C2 + i] = z;
        }
        __syncthreads();
        src_base = k_lo - static_cast<int>(C2);
        new_lo = k_lo, new_hi = k_hi;
    }
    // 2. the windows whose last element is one the tile added: u*S + W - 1 in [new_lo, new_hi]
    int u_lo = new_lo - static_cast<int>(W - 1);
    u_lo = u_lo <= 0 ? 0 : (u_lo + static_cast<int>(S) - 1) / static_cast<int>(S);
    int u_hi = new_hi - static_cast<int>(W - 1); // inclusive
    u_hi = u_hi < 0 ? -1 : u_hi / static_cast<int>(S);
    if (u_hi >= static_cast<int>(a.tail_units)) u_hi = static_cast<int>(a.tail_units) - 1;
    const uint32_t nwin = u_hi >= u_lo ? static_cast<uint32_t>(u_hi - u_lo + 1) : 0u;
    if (!warm && nwin) {
        // leaves
        for (uint32_t i = tid; i < nwin * W; i += NT) {
            const uint32_t w = i >> logw, n = i & (W - 1);
            const uint32_t y = static_cast<uint32_t>((u_lo + static_cast<int>(w)) * static_cast<int>(S) + static_cast<int>(n) - src_base);
            Wk[(w << logw) + leaf_position(n, W, n_r4, odd)] = from_y ? ((n >= W - T) ? Ys[ysk(y)] : Yv[ysk(y)]) : Zv[y];
        }
        __syncthreads();
        if (odd) {
            for (uint32_t b = tid; b < nwin * (W / 2); b += NT) {
                const float2 p = Wk[2 * b], q = Wk[2 * b + 1];
                Wk[2 * b] = padd(p, q);
                Wk[2 * b + 1] = psub(p, q);
            }
            __syncthreads();
        }
        int logq = odd ? 1 : 0;
        for (uint32_t q = odd ? 2 : 1; q < W; q <<= 2, logq += 2) {
            const uint32_t scale = W >> (logq + 2);
            for (uint32_t b = tid; b < nwin * (W / 4); b += NT) {
                const uint32_t wdw = b >> (logw - 2), bb = b & ((W >> 2) - 1);
                const uint32_t blk = bb >> logq, k = bb & (q - 1);
                float2 *base = Wk + (wdw << logw) + (blk << (logq + 2)) + k;
                float2 t0 = base[0], t1 = base[q], t2 = base[2 * q], t3 = base[3 * q];
                if (k != 0) {
                    t1 = pmul_tw(t1, __ldg(a.fft.tw + k * scale), a.one);
                    t2 = pmul_tw(t2, __ldg(a.fft.tw + 2 * k * scale), a.one);
                    t3 = pmul_tw(t3, __ldg(a.fft.tw + 3 * k * scale), a.one);
                }
                pradix4(t0, t1, t2, t3);
                base[0] = t0;
                base[q] = t1;
                base[2 * q] = t2;
                base[3 * q] = t3;
            }
            __syncthreads();
        }
        // epilogue: four consecutive bins per step leave as one word when the output is index-only
        const bool words = W >= 8 && !a.fft.mag && a.fft.use_thr && (reinterpret_cast<uintptr_t>(a.fft.idx) & 3) == 0;
        if (words) {
            for (uint32_t i = 4 * tid; i < nwin * W; i += 4 * NT) {
                const uint32_t w = i >> logw, pos = i & (W - 1);
                *reinterpret_cast<uint32_t *>(a.fft.idx + static_cast<uint64_t>(u_lo + w) * W + ((pos + W / 2) & (W - 1))) =
                    glyph4_word(a.fft, Wk[i], Wk[i + 1], Wk[i + 2], Wk[i + 3]);
            }
        } else {
            for (uint32_t i = tid; i < nwin * W; i += NT) emit_bin(a.fft, static_cast<uint64_t>(u_lo + (i >> logw)), W, i & (W - 1), Wk[i]);
        }
    }
    // 3. the next tile's carries: the last C1 stream outputs (a partial last tile has no successor) and the last C2
    // second-stage outputs
    for (uint32_t i = tid; i < C1; i += NT) {
        Cv[i] = Yv[ysk(T_TILE + i)];
        Cs[i] = Ys[ysk(T_TILE + i)];
    }
    for (uint32_t i = tid; i < C2; i += NT) Cz[i] = Zv[n2 + i];
}

// resident CTAs per SM the kernel is compiled for (register budget) and launched at
template <int D, int R, int NT, bool EXACT, int LS>
constexpr int ctas_per_sm()
{
    if (NT <= 128 && D <= 8) {
        if (LS > 0) return (R <= 2 ? 8 : (EXACT ? QD_EXP_EXACT_CTAS : 5)) * (128 / NT);
        return 4 * (128 / NT);
    }
    return NT <= 128 ? 3 : 2; // long-filter shapes: a 74 KB sample layout, tiles read from global memory (no staging buffer)
}

// SNAP: the launch also leaves snapshots of running sums (FirArgs::tail_out); a kernel of its own, so that each
// kernel holds ONE copy of the filter loops (with two copies in one kernel ptxas keeps the tap loads of one of
// them out of the uniform registers)
// FUSE: the launch's units are back-to-back sparkfft windows and the STFT runs in the same kernel (FirArgs::fft)
template <int D, int R, int NT, bool EXACT, int LS, bool SNAP = false, int FUSE = 0>
__global__ void __launch_bounds__(NT, (ctas_per_sm<D, R, NT, EXACT, LS>())) fk_fir(const __grid_constant__ FirArgs a, const __grid_constant__ FirTaps taps)
{
    constexpr int LMAX = LS > 0 ? LS : kMaxTapPairs;
    using Gm = FirGeom<D, R, NT, LMAX>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    LeanPhase *lphase = reinterpret_cast<LeanPhase *>(smem + 16);
    float2 *X = reinterpret_cast<float2 *>(smem + kSmemHeader);
    uint8_t *raw0 = smem + kSmemHeader + Gm::X_BYTES;
    double2 *ttab = reinterpret_cast<double2 *>(raw0 + a.raw_cap); // lean path: e^{i 4t ratio} per thread
    const int tid = threadIdx.x;
    // FAST, cs8, at most one shift: the lean decode loop (tiles that start off a 4-sample boundary or straddle a
    // binade of n*ratio take the general one)
    const bool lean = !EXACT && a.fmt == QD_FMT_CS8 && a.n_shift <= 1;
    const bool lean_mix = lean && a.n_shift == 1;
    const uint32_t pb = a.fmt == QD_FMT_CF32 ? 8 : (a.fmt == QD_FMT_CS16 ? 4 : 2);
    // cf32 tiles are read straight from global memory; so are integer tiles when the host gave no staging buffer
    // (EXACT mode with a long-filter shape: the buffer would cost a resident CTA)
    const bool staged = a.fmt != QD_FMT_CF32 && a.raw_cap != 0;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto tile_phase = [&](const TileGeo &g) { // one spare lane, one tile ahead
        // the span of a FULL tile, whatever part of it this launch wants: `ok` must not depend on the launch
        lean_phase(a, g.n_tile0, static_cast<uint32_t>(Gm::T_TILE - 1) * D + a.L + 4, lphase);
    };
    if (lean_mix) {
        double c, s;
        sincos_f64k(__dmul_rn(static_cast<double>(4 * tid), a.ratio[0]), a.sincos, a.k, c, s);
        ttab[tid] = make_double2(c, s);
        if (tid == 0 && blockIdx.x < a.n_tiles) tile_phase(tile_geo<D, Gm::T_TILE>(a, blockIdx.x));
    }
    __syncthreads();

    // raw byte range of a tile, widened to 16-byte boundaries for the bulk copy.  Local sample 0 (n_tile0) of a
    // launch's first tile may lie before the launch's first wanted sample -- even before the resident range --
    // so addresses are formed in integer arithmetic and only [l_lo, n_dec) is touched.
    auto issue = [&](const TileGeo &g) {
        const uint64_t span = static_cast<uint64_t>(g.cnt - 1) * D + a.L;
        const uint64_t n_dec = min(span, a.src_end - g.n_tile0);
        const uintptr_t g0 = reinterpret_cast<uintptr_t>(a.src) + (g.n_tile0 - a.src_base) * pb; // wraps consistently
        const uintptr_t org = g0 & ~uintptr_t(15);
        const uintptr_t abeg = (g0 + static_cast<uint64_t>(g.skip) * D * pb) & ~uintptr_t(15);
        const uintptr_t aend = (g0 + n_dec * pb + 15) & ~uintptr_t(15);
        const uint32_t bytes = static_cast<uint32_t>(aend - abeg);
        if (staged) {
            mbar_expect_tx(&mbar[0], bytes);
            bulk_g2s(raw0 + (abeg - org), reinterpret_cast<const void *>(abeg), bytes, &mbar[0]);
        } else {
            bulk_prefetch_l2(reinterpret_cast<const void *>(abeg), bytes); // cf32: the decode stage reads global memory; make it an L2 hit
        }
    };

    // Tiles of a CTA: every gridDim.x-th one -- or, when windows are carried from tile to tile (FUSE == 2), a
    // contiguous run preceded by one warm-up tile that only fills the carry
    uint64_t t_first = blockIdx.x, t_begin = blockIdx.x, t_last = a.n_tiles, t_step = gridDim.x;
    if constexpr (FUSE == 2) {
        const uint64_t per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
        t_first = blockIdx.x * per;
        t_last = min(t_first + per, a.n_tiles);
        t_begin = t_first > 0 ? t_first - 1 : 0;
        t_step = 1;
        if (t_first >= t_last) t_begin = t_last; // no tile for this CTA
    }
    // CARRY (long run-time-length filters over unstaged tiles -- integer formats, cf32 behind a shift: configs 4 and 1):
    // consecutive tiles overlap
    // by L - D samples, a tenth of a tile at L = 800, D = 16, and decode + f64 mixer are 40 % of such a tile's time.
    // A CTA therefore walks a contiguous run of tiles and moves the overlap -- the last L - D decoded and mixed
    // samples, which are the next tile's first -- across in registers (the same row of the polyphase layout,
    // T_TILE*D/DR columns to the left) instead of decoding and mixing it again.  Values are pure functions of the
    // absolute sample index, so nothing changes bit for bit.
    constexpr bool CARRY = EXACT && LS == 0 && FUSE <= 1;
    constexpr uint32_t CARRY_MAX = 4; // float4 per thread held across the STFT
    constexpr uint32_t SHIFT_COLS = static_cast<uint32_t>(Gm::T_TILE) * D / Gm::DR;
    const uint32_t ov_groups = (a.L > static_cast<uint32_t>(D) && ((a.L - D) & 3) == 0) ? (a.L - D) / 4 : 0;
    const uint32_t span_full = static_cast<uint32_t>(Gm::T_TILE - 1) * D + a.L;
    const bool carry_run = CARRY && !a.no_carry && !staged && (a.fmt != QD_FMT_CF32 || a.n_shift != 0) && a.contiguous && ov_groups != 0 &&
                           2 * ov_groups <= CARRY_MAX * NT && (Gm::T_TILE * D) % Gm::DR == 0 && a.n_tiles > gridDim.x;
    if (CARRY && carry_run) {
        const uint64_t per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
        t_first = t_begin = blockIdx.x * per;
        t_last = min(t_first + per, a.n_tiles);
        t_step = 1;
        if (t_first >= t_last) t_begin = t_last; // no tile for this CTA
    }
    bool have_carry = false;
    uint64_t it = 0;
    if (tid == 0 && t_begin < t_last) issue(tile_geo<D, Gm::T_TILE>(a, t_begin));

    for (uint64_t tile = t_begin; tile < t_last; tile += t_step, ++it) {
        const TileGeo g = tile_geo<D, Gm::T_TILE>(a, tile);
        const uint64_t span = static_cast<uint64_t>(g.cnt - 1) * D + a.L;
        const uint32_t n_dec = static_cast<uint32_t>(min(span, a.src_end - g.n_tile0));
        const uintptr_t gbeg = reinterpret_cast<uintptr_t>(a.src) + (g.n_tile0 - a.src_base) * pb;
        const uint32_t lead = static_cast<uint32_t>(gbeg & 15) / pb;
        const uint32_t l_lo = g.skip * D; // first wanted local sample (0 except in a launch's first tile)

        if (staged) mbar_wait(&mbar[0], static_cast<uint32_t>(it & 1));

        // ---- decode + mix once per sample, into the polyphase layout ------------------------------
        bool decoded_lean = false; // through decode_exact_global from local sample 0 on: the whole span is in X
        {
            const uint8_t *raw = staged ? raw0 : reinterpret_cast<const uint8_t *>(gbeg & ~uintptr_t(15));
            if (!EXACT && lean && (lead & 3) == 0 && (!lean_mix || lphase->ok)) {
                if (lean_mix) decode_tile_lean<D, R, NT, LMAX, true>(a, raw, lead, l_lo, n_dec, g.n_tile0, lphase, ttab, X, tid);
                else decode_tile_lean<D, R, NT, LMAX, false>(a, raw, lead, l_lo, n_dec, g.n_tile0, lphase, ttab, X, tid);
            } else if (a.fmt == QD_FMT_CF32 && a.n_shift == 0 && lead == 0 && l_lo == 0) {
                decode_cf32_copy<Gm, NT>(raw, n_dec, X, tid);
            } else if (EXACT && a.fmt == QD_FMT_CF32 && lead == 0 && l_lo == 0) {
                // cf32 behind shifts: the same lean loop as the unstaged integer tiles (packed mixer, next group in flight)
                decode_exact_global<Gm, NT, QD_FMT_CF32>(a, raw, n_dec, g.n_tile0, reinterpret_cast<float4 *>(X), tid,
                                                         (CARRY && have_carry) ? ov_groups : 0u);
                decoded_lean = true;
            } else if (EXACT && !staged && a.fmt != QD_FMT_CF32 && (lead & 3) == 0 && l_lo == 0) {
                const uint8_t *g0 = raw + pb * lead;
                float4 *X4 = reinterpret_cast<float4 *>(X);
                const uint32_t gf = (CARRY && have_carry) ? ov_groups : 0u; // the overlap is already in place
                switch (a.fmt) {
                case QD_FMT_CS8: decode_exact_global<Gm, NT, QD_FMT_CS8>(a, g0, n_dec, g.n_tile0, X4, tid, gf); break;
                case QD_FMT_CU8: decode_exact_global<Gm, NT, QD_FMT_CU8>(a, g0, n_dec, g.n_tile0, X4, tid, gf); break;
                default: decode_exact_global<Gm, NT, QD_FMT_CS16>(a, g0, n_dec, g.n_tile0, X4, tid, gf); break;
                }
                decoded_lean = true;
            } else if (EXACT && staged && (lead & 3) == 0 && l_lo == 0) {
                const uint32_t raw_addr = smem_u32(raw) + pb * lead;
                float4 *X4 = reinterpret_cast<float4 *>(X);
                switch (a.fmt) {
                case QD_FMT_CS8: decode_lean_exact<Gm, NT, QD_FMT_CS8>(a, raw_addr, n_dec, g.n_tile0, X4, tid); break;
                case QD_FMT_CU8: decode_lean_exact<Gm, NT, QD_FMT_CU8>(a, raw_addr, n_dec, g.n_tile0, X4, tid); break;
                default: decode_lean_exact<Gm, NT, QD_FMT_CS16>(a, raw_addr, n_dec, g.n_tile0, X4, tid); break;
                }
            } else if ((lead & 3) == 0) {
                switch (a.fmt) {
                case QD_FMT_CS8: decode_tile<QD_FMT_CS8, D, R, NT, LMAX, true, !EXACT>(a, raw, lead, l_lo, n_dec, g.n_tile0, X, tid); break;
                case QD_FMT_CU8: decode_tile<QD_FMT_CU8, D, R, NT, LMAX, true, !EXACT>(a, raw, lead, l_lo, n_dec, g.n_tile0, X, tid); break;
                case QD_FMT_CS16: decode_tile<QD_FMT_CS16, D, R, NT, LMAX, true, !EXACT>(a, raw, lead, l_lo, n_dec, g.n_tile0, X, tid); break;
                default: decode_tile<QD_FMT_CF32, D, R, NT, LMAX, true, !EXACT>(a, raw, lead, l_lo, n_dec, g.n_tile0, X, tid); break;
                }
            } else {
                switch (a.fmt) {
                case QD_FMT_CS8: decode_tile<QD_FMT_CS8, D, R, NT, LMAX, false, !EXACT>(a, raw, lead, l_lo, n_dec, g.n_tile0, X, tid); break;
                case QD_FMT_CU8: decode_tile<QD_FMT_CU8, D, R, NT, LMAX, false, !EXACT>(a, raw, lead, l_lo, n_dec, g.n_tile0, X, tid); break;
                case QD_FMT_CS16: decode_tile<QD_FMT_CS16, D, R, NT, LMAX, false, !EXACT>(a, raw, lead, l_lo, n_dec, g.n_tile0, X, tid); break;
                default: decode_tile<QD_FMT_CF32, D, R, NT, LMAX, false, !EXACT>(a, raw, lead, l_lo, n_dec, g.n_tile0, X, tid); break;
                }
            }
        }
        __syncthreads();
        // the raw bytes are consumed: fetch the next tile's while this one is filtered
        if (tid == 0 && tile + t_step < t_last) issue(tile_geo<D, Gm::T_TILE>(a, tile + t_step));
        // the next tile's phase state: another warp's spare lane, so no warp carries both chores into the barrier
        if (lean_mix && tid == (NT > 32 ? 32 : 0) && tile + t_step < t_last)
            tile_phase(tile_geo<D, Gm::T_TILE>(a, tile + t_step));

        // ---- FIR: thread owns outputs R*tid .. R*tid+R-1 of the tile ------------------------------
        float2 acc[R], snapv[R];
        uint32_t snap_mask;
        const bool mine = fir_tile<D, R, NT, LMAX, EXACT, LS, SNAP, FUSE>(a, taps, g, X, tid, tid, acc, snapv, snap_mask);
        __syncthreads();
        // the overlap with the CTA's next tile leaves X before the STFT reuses its first bytes
        float4 cr[CARRY_MAX];
        const bool carry_next = CARRY && carry_run && decoded_lean && n_dec == span_full && tile + 1 < t_last;
        auto carry_slot = [&](uint32_t e) { // element e of the overlap: group e (first halves), then group e - ov_groups (second halves)
            const uint32_t half = e >= ov_groups ? 1u : 0u, gq = e - half * ov_groups;
            return ((gq & (Gm::G - 1)) + half * Gm::G) * Gm::PITCH + (gq >> Gm::LOG_G);
        };
        if (CARRY && carry_next) {
            const float4 *X4 = reinterpret_cast<const float4 *>(X);
#pragma unroll
            for (uint32_t j = 0; j < CARRY_MAX; j++) {
                const uint32_t e = tid + j * NT;
                if (e < 2 * ov_groups) cr[j] = X4[carry_slot(e) + SHIFT_COLS];
            }
        }
        if constexpr (FUSE == 1) {
            fused_stft<R, NT, Gm::T_TILE>(a, g, X, acc, mine, tid); // X is free: every thread has left the filter
            __syncthreads();
        }
        if (CARRY && carry_next) { // disjoint from what the next decode writes (groups >= ov_groups); it ends with a barrier
            float4 *X4 = reinterpret_cast<float4 *>(X);
#pragma unroll
            for (uint32_t j = 0; j < CARRY_MAX; j++) {
                const uint32_t e = tid + j * NT;
                if (e < 2 * ov_groups) X4[carry_slot(e)] = cr[j];
            }
        }
        have_carry = carry_next;
        if constexpr (FUSE == 2) {
            fused_stream_stft<R, NT, Gm::T_TILE>(a, g, X, reinterpret_cast<float2 *>(smem + a.carry_off), acc, snapv, snap_mask, mine,
                                                 tile < t_first, tid);
            __syncthreads();
        }
    }
}

template <int D, int R, int NT, bool EXACT, int LS>
static int launch_fir_k(Chain &c, const FirArgs &a, const FirTaps &t)
{
    using Gm = FirGeom<D, R, NT, (LS > 0 ? LS : kMaxTapPairs)>;
    size_t smem = kSmemHeader + Gm::X_BYTES + static_cast<size_t>(a.raw_cap) + (EXACT ? 0 : NT * sizeof(double2));
    FirArgs a2 = a;
    const bool stream_fuse = a.fft.W && a.fuse_S; // overlapping windows carried between tiles (FUSE = 2)
    if (stream_fuse) {
        // scratch inside the sample layout: values + snapshots of carry and tile, one work array per window
        const size_t C1 = a.st2_L ? a.st2_L - 1 : a.fft.W - 1, C2 = a.st2_L ? a.fft.W - 1 : 0;
        const size_t n2 = a.st2_L ? Gm::T_TILE / a.st2_D + 2 : 0;
        const size_t nwin_max = (a.st2_L ? n2 : Gm::T_TILE) / a.fuse_S + 2;
        const size_t yn = C1 + Gm::T_TILE + ((C1 + Gm::T_TILE) >> 5) + 1; // skewed (fused_stream_stft ysk)
        if (a.total_out >= (uint64_t(1) << 31) || a.tail_units >= (uint64_t(1) << 31) ||
            (2 * yn + C2 + n2 + nwin_max * a.fft.W) * sizeof(float2) > Gm::X_BYTES)
            return set_error(QD_E_INVALID_ARG, "internal: fused stream STFT scratch does not fit the sample layout");
        a2.carry_off = static_cast<uint32_t>((smem + 15) / 16 * 16);
        smem = a2.carry_off + (2 * C1 + C2) * sizeof(float2);
    }
    if (smem > 227 * 1024) return set_error(QD_E_INVALID_ARG, "internal: fused FIR tile needs %zu bytes of shared memory", smem);
    int per_sm = std::max<int>(1, static_cast<int>((227 * 1024) / (smem + 1024)));
    if (c.fir_cta_cap > 0) per_sm = std::min(per_sm, c.fir_cta_cap);
    const int grid = static_cast<int>(std::min<uint64_t>(a.n_tiles, static_cast<uint64_t>(c.ctx->sm_count) * std::min(per_sm, ctas_per_sm<D, R, NT, EXACT, LS>())));
    if (stream_fuse) {
        if constexpr (EXACT) {
            QD_CUDA(cudaFuncSetAttribute(fk_fir<D, R, NT, EXACT, LS, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            fk_fir<D, R, NT, EXACT, LS, true, 2><<<grid, NT, smem, c.stream>>>(a2, t);
        } else {
            return set_error(QD_E_INVALID_ARG, "internal: the fused sparkfft kernels are EXACT only");
        }
    } else if (a.fft.W) {
        // instantiated for the run-time-length filter in EXACT arithmetic (the sparkfft sink needs bit-exact indices)
        if constexpr (LS == 0 && EXACT) {
            QD_CUDA(cudaFuncSetAttribute(fk_fir<D, R, NT, EXACT, LS, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            fk_fir<D, R, NT, EXACT, LS, false, 1><<<grid, NT, smem, c.stream>>>(a, t);
        } else {
            return set_error(QD_E_INVALID_ARG, "internal: no fused sparkfft kernel for this filter shape");
        }
    } else if (a.tail_out) {
        QD_CUDA(cudaFuncSetAttribute(fk_fir<D, R, NT, EXACT, LS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        fk_fir<D, R, NT, EXACT, LS, true><<<grid, NT, smem, c.stream>>>(a, t);
    } else {
        QD_CUDA(cudaFuncSetAttribute(fk_fir<D, R, NT, EXACT, LS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        fk_fir<D, R, NT, EXACT, LS, false><<<grid, NT, smem, c.stream>>>(a, t);
    }
    QD_LAUNCHED();
    return QD_OK;
}

// LS = 40 is the reference's default filter (args.rs:165: `None => 40`), specialised at compile time
template <int D, int R, int NT>
static int launch_fir_dr(Chain &c, const FirArgs &a, const FirTaps &t, bool exact)
{
    if (a.L == 40) return exact ? launch_fir_k<D, R, NT, true, 40>(c, a, t) : launch_fir_k<D, R, NT, false, 40>(c, a, t);
    return exact ? launch_fir_k<D, R, NT, true, 0>(c, a, t) : launch_fir_k<D, R, NT, false, 0>(c, a, t);
}


// one translation unit per decimation (qd_fast_dN.cu)
int launch_fir_d2(Chain &c, const FirArgs &a, const FirTaps &t, bool exact);
int launch_fir_d4(Chain &c, const FirArgs &a, const FirTaps &t, bool exact);
int launch_fir_d8(Chain &c, const FirArgs &a, const FirTaps &t, bool exact);
int launch_fir_d16(Chain &c, const FirArgs &a, const FirTaps &t, bool exact);
int launch_fir_d32(Chain &c, const FirArgs &a, const FirTaps &t, bool exact);

} // namespace qd
