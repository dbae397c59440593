// qd_fast.cu -- the fused hot-path kernel: decode + NCO mix + decimating FIR in one pass.
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
#include <algorithm>
#include <cmath>
#include <cstring>

#include <unistd.h>

#include "qd_fir_kernel.cuh"
#include "qd_tcfir.h"

namespace qd {

// ---------------------------------------------------------------------------- truncated window tails
// A few windows per warp: output k = n - T + r of the read (off, n) is the one whose taps stop at the end of that
// read's raw buffer, J = (T - r) * D + L/2 < L of them (filter.rs:68-71,107-124).  The T tails of a window
// share their T*D + L/2 samples, so the warp decodes and mixes those once into shared memory, then lane r sums
// tail r in ascending tap order.  Always the exact arithmetic: decode, fl64(n * ratio) and an f64 sin/cos per
// sample and shift, product and sum rounded separately.
struct TailArgs {
    const uint8_t *src;
    uint64_t src_base;
    int fmt, n_shift;
    double ratio[kMaxLeadShifts];
    const double *sincos;
    uint32_t L, D, T, span; // span = T*D + L/2 samples per window
    uint32_t log_d;         // D = 2^log_d
    uint32_t wpw;           // windows per warp: min(32 / T, 8)
    uint64_t off0, S, n_call, n_units;
    float2 *out; // tail r of window u at out[u * out_pitch + out_off + r]: a patch matrix [n_units][T] (pitch T, offset 0)
                 // or the tails' own places in a [n_units][n_call] output (pitch n_call, offset n_call - T)
    uint64_t out_pitch, out_off;
    float2 one;
};
constexpr int kTailWarps = 4;

__global__ void __launch_bounds__(32 * kTailWarps) fk_tail(const __grid_constant__ TailArgs a, const __grid_constant__ FirTaps taps)
{
    extern __shared__ float2 tail_smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a warp takes wpw consecutive windows, so that wpw * T of its lanes run summation loops
    const uint64_t u0 = (static_cast<uint64_t>(blockIdx.x) * kTailWarps + w) * a.wpw;
    if (u0 >= a.n_units) return;
    const uint32_t nw = static_cast<uint32_t>(min(static_cast<uint64_t>(a.wpw), a.n_units - u0));
    // sample l of a window sits at l + l/D: the lanes of the summation loop read D samples apart, which without
    // the skew is one shared-memory bank
    const uint32_t pitch = a.span + (a.span >> a.log_d) + 1;
    float2 *x = tail_smem + static_cast<size_t>(w) * a.wpw * pitch;
    const uint32_t i0 = a.L - a.L / 2;
    // all nw * span samples of the warp's windows in one flat loop (a window's span is often shorter than a warp)
    const uint64_t n00 = (a.off0 + u0 * a.S + a.n_call - a.T) * a.D + i0; // sample under tap 0 of window u0's first tail
    const uint64_t wstep = a.S * a.D;                                      // raw samples from one window to the next
    {
        uint32_t g = 0, l = lane;
        for (uint32_t idx = lane; idx < nw * a.span; idx += 32, l += 32) {
            while (l >= a.span) {
                l -= a.span;
                g++;
            }
            const uint64_t n = n00 + g * wstep + l;
            float2 v = decode_sample(a.src, a.fmt, n - a.src_base);
            for (int sft = 0; sft < a.n_shift; sft++) v = cmul_exact(v, phasor_exact(n, a.ratio[sft], a.sincos));
            x[g * pitch + l + (l >> a.log_d)] = v;
        }
    }
    __syncwarp();
    const uint32_t g = static_cast<uint32_t>(lane) / a.T, r = static_cast<uint32_t>(lane) % a.T;
    if (g < nw && a.T <= 32) {
        const uint32_t J = min(a.L, (a.T - r) * a.D + a.L / 2);
        const float2 *xg = x + g * pitch;
        float2 acc = make_float2(0.0f, 0.0f);
        for (uint32_t j = 0, l = r * a.D; j < J; j++, l++) acc = fma2(mul2(xg[l + (l >> a.log_d)], taps.t[j]), a.one, acc); // fl(acc + fl(x * f)), filter.rs:119
        a.out[(u0 + g) * a.out_pitch + a.out_off + r] = acc;
    }
}

// ---------------------------------------------------------------------------- host side

// (decimate) -> outputs per thread R and threads per CTA; R * NT * D ~ 8192 raw samples per tile
struct FirShape {
    int R, NT;
};
static bool fir_shape(uint64_t D, FirShape *s)
{
    switch (D) {
    case 2: *s = {8, 128}; return true;
    case 4: *s = {8, 128}; return true;
    case 8: *s = {4, 128}; return true;
    case 16: *s = {4, 128}; return true;
    case 32: *s = {2, 128}; return true;
    }
    return false;
}

// One LowPass stage as the fused kernel sees it
struct LpInfo {
    const Stage *st = nullptr;
    uint32_t D = 0, L = 0, i0 = 0;
    uint32_t T = 0; // outputs at the end of every read that use a truncated filter: (n-k)*D + L/2 < L
    FirShape shape{0, 0};
};

struct FastPlan {
    bool ok = false;
    int n_shift = 0;
    int n_lp = 0;
    LpInfo lp[2];
    // The top stage's outputs do not depend on the read they belong to (no truncated positions), so it is
    // materialised once as a contiguous stream and windows are taken from it at the sink's stride.
    bool stream_top = false;
    // Overlapping windows behind ONE filter that does have truncated positions (T > 0): all but the last T
    // samples of a window are still read-independent, so the stream is shared as above and the last T samples of
    // every window -- the ones whose taps stop at the end of that window's own raw buffer -- are computed on
    // their own (fk_tail) into a patch matrix the STFT kernel reads in their place.
    bool stream_tail = false;
};

static bool lp_info(const Stage &st, LpInfo *o)
{
    if (st.kind != QD_STAGE_LOWPASS || !fir_shape(st.decimate, &o->shape)) return false;
    o->st = &st;
    o->D = static_cast<uint32_t>(st.decimate);
    o->L = static_cast<uint32_t>(st.size);
    const uint64_t Q = (st.size + o->D - 1) / o->D;
    if (Q * o->D > static_cast<uint64_t>(kMaxTapPairs)) return false;
    o->i0 = o->L - o->L / 2;
    o->T = (o->i0 + o->D - 1) / o->D - 1;
    return true;
}

static FastPlan fast_plan(const Chain &c, uint64_t unit_len, uint64_t stride, uint64_t n_units)
{
    FastPlan f;
    const Source &s = c.src;
    if (s.kind == QD_SRC_GEN) return f;
    const size_t S = c.stages.size();
    size_t i = 0;
    while (i < S && c.stages[i].kind == QD_STAGE_SHIFT) i++;
    if (i > static_cast<size_t>(kMaxLeadShifts)) return f;
    f.n_shift = static_cast<int>(i);
    f.n_lp = static_cast<int>(S - i);
    if (f.n_lp < 1 || f.n_lp > 2) return f;
    for (int k = 0; k < f.n_lp; k++)
        if (!lp_info(c.stages[i + k], &f.lp[k])) return f;
    const LpInfo &top = f.lp[f.n_lp - 1];
    if (f.n_lp == 2) {
        // The inner stage can only be shared between units if no unit ever reads the truncated tail of
        // its own inner read: the outer stage must be free of truncation (T == 0, i.e. D2 >= i0_2) and its
        // last tap must stop T1 samples short of the inner buffer's end (D2 - i0_2 >= T1).
        if (top.T != 0 || top.D - top.i0 < f.lp[0].T) return f;
        f.stream_top = true;
    } else {
        f.stream_top = top.T == 0 && stride != unit_len && n_units > 1;
        // windows further apart than they are wide: a stream would also compute (and, for host sources, budget
        // for) the gaps between them; evaluate the windows on their own when the per-unit kernel can
        if (f.stream_top && stride > unit_len && unit_len % static_cast<uint64_t>(top.shape.R) == 0) f.stream_top = false;
        if (!f.stream_top && c.allow_tail && top.T > 0 && top.T <= 32 && top.T < unit_len && stride < unit_len && n_units > 1) {
            f.stream_top = true;
            f.stream_tail = true;
        }
    }
    if (!f.stream_top && unit_len % static_cast<uint64_t>(top.shape.R) != 0) return f;
    // absolute sample 0 must sit on a 16-byte boundary so every tile's bytes can be bulk-copied
    if (s.kind == QD_SRC_DEVICE_MEM) {
        const uint64_t pb = pair_bytes(s.format);
        if ((reinterpret_cast<uintptr_t>(s.data) - static_cast<uintptr_t>(s.base_sample * pb)) % 16 != 0) return f;
    }
    f.ok = true;
    return f;
}

// FUSE = 2 scratch (values + snapshots of carry and tile, one work array per window of a tile) inside the sample
// layout of the kernel launch_fir_dr picks for this stage: the same arithmetic as launch_fir_k
static bool fuse_scratch_fits(const LpInfo &lp, uint64_t W, uint64_t S)
{
    const int D = static_cast<int>(lp.D), R = lp.shape.R, NT = lp.shape.NT;
    const int ls = lp.L == 40 ? 40 : 0, lmax = ls ? ls : kMaxTapPairs;
    const uint64_t t_out = static_cast<uint64_t>(R) * NT, t_tile = static_cast<uint64_t>(tile_outputs(D, R, NT, ls));
    const uint64_t span = (t_out - 1) * D + static_cast<uint64_t>((lmax + D - 1) / D) * D + ((lmax % 4 || D % 4) ? 3 : 0);
    const uint64_t dr = static_cast<uint64_t>(D) * R, cols = (span + dr - 1) / dr;
    const uint64_t x_bytes = (dr / 2) * static_cast<uint64_t>(pitch_for(static_cast<int>(dr / 4), static_cast<int>(cols))) * sizeof(float4);
    const uint64_t yn = W - 1 + t_tile + ((W - 1 + t_tile) >> 5) + 1;
    const uint64_t need = (2 * yn + (t_tile / S + 2) * W) * sizeof(float2);
    return need <= x_bytes;
}

// `lowpass | lowpass | sparkfft` in one kernel (FirArgs::st2_L): the outer filter's taps fit the argument block, a
// tile completes enough outer outputs for a chunk's warm-up tile to fill the window carry, and the scratch fits
static bool fuse_two_stage_fits(const LpInfo &in, const LpInfo &top, uint64_t W, uint64_t S)
{
    if (top.L > 256 || top.T != 0) return false;
    const int D = static_cast<int>(in.D), R = in.shape.R, NT = in.shape.NT;
    const int ls = in.L == 40 ? 40 : 0, lmax = ls ? ls : kMaxTapPairs;
    const uint64_t t_out = static_cast<uint64_t>(R) * NT, t_tile = static_cast<uint64_t>(tile_outputs(D, R, NT, ls));
    if (top.L - 1 > t_tile) return false;
    // outer outputs a tile completes, less those of a warm-up tile that would need the tile before it
    if (t_tile / top.D < (top.L - 1 + top.D - 1) / top.D + (W - 1) + 1) return false;
    const uint64_t span = (t_out - 1) * D + static_cast<uint64_t>((lmax + D - 1) / D) * D + ((lmax % 4 || D % 4) ? 3 : 0);
    const uint64_t dr = static_cast<uint64_t>(D) * R, cols = (span + dr - 1) / dr;
    const uint64_t x_bytes = (dr / 2) * static_cast<uint64_t>(pitch_for(static_cast<int>(dr / 4), static_cast<int>(cols))) * sizeof(float4);
    const uint64_t n2 = t_tile / top.D + 2;
    const uint64_t yn = top.L - 1 + t_tile + ((top.L - 1 + t_tile) >> 5) + 1;
    const uint64_t need = (2 * yn + (W - 1) + n2 + (n2 / S + 2) * W) * sizeof(float2);
    return need <= x_bytes;
}

// One fused-kernel launch.  The source is raw capture bytes (fmt, shifts) or a cf32 stream from an earlier
// stage.  n_units units of n_call outputs at unit stride S (top-level samples) starting at off0, written as
// [n_units][n_call]; total_out limits the count in contiguous mode.
// truncated window tails taken as snapshots inside a stream launch (see FirArgs::tail_out)
struct TailSnap {
    float2 *out;
    uint32_t W, S, T;
    uint64_t units;
};

static int launch_fir(Chain &c, const LpInfo &lp, int fmt, int n_shift, const double *ratios, const uint8_t *d_src,
                      uint64_t src_base, uint64_t src_end, uint64_t off0, uint64_t n_call, uint64_t S, uint64_t n_units,
                      uint64_t total_out, float2 *d_out, const TailSnap *snap = nullptr, const FftArgs *fuse = nullptr,
                      uint32_t fuse_stride = 0, const LpInfo *stage2 = nullptr, uint64_t stage2_total = 0)
{
    const Stage &st = *lp.st;
    FirArgs a;
    memset(&a, 0, sizeof a);
    a.src = d_src;
    a.src_base = src_base;
    a.src_end = src_end;
    a.fmt = fmt;
    a.n_shift = n_shift;
    for (int i = 0; i < n_shift; i++) a.ratio[i] = ratios[i];
    a.sincos = c.ctx->d_sincos;
    a.k = make_sincos_k();
    const uint32_t D = lp.D, R = static_cast<uint32_t>(lp.shape.R);
    a.L = lp.L;
    a.off0 = off0;
    a.n_call = n_call;
    a.S = S;
    a.n_units = n_units;
    // units that tile the output stream are filtered as ONE stream cut into absolute tiles; a thread's R outputs
    // must then never straddle a unit boundary (off0 and n_call multiples of R)
    a.contiguous = (n_units == 1 || (S == n_call && off0 % R == 0 && n_call % R == 0)) ? 1 : 0;
    a.no_carry = c.fir_carry ? 0 : 1;
    a.ncall_log2 = -1;
    if (is_pow2(n_call))
        for (int b = 0; b < 64; b++)
            if ((uint64_t(1) << b) == n_call) a.ncall_log2 = b;
    a.total_out = total_out;
    const uint64_t t_out = static_cast<uint64_t>(tile_outputs(static_cast<int>(D), static_cast<int>(R), lp.shape.NT, a.L == 40 ? 40 : 0)); // as FirGeom::T_TILE of the kernel launch_fir_dr picks
    a.tiles_per_unit = static_cast<uint32_t>((n_call + t_out - 1) / t_out);
    if (a.contiguous) { // absolute tile numbering: see FirArgs::tile_first
        a.tile_first = off0 / t_out;
        a.n_tiles = total_out ? (off0 + total_out - 1) / t_out - a.tile_first + 1 : 0;
    } else {
        a.n_tiles = n_units * a.tiles_per_unit;
    }
    const uint64_t pb = pair_bytes(fmt);
    const uint64_t span_max = (t_out - 1) * D + a.L;
    // integer tiles are staged in shared memory by bulk copy, except in EXACT mode with the long-filter shapes
    // (D >= 16): there the staging buffer would cost a resident CTA and the tile is read from global memory
    const bool exact_mode = c.precision == QD_PRECISION_EXACT;
    const bool unstaged = fmt == QD_FMT_CF32 || (exact_mode && D >= 16);
    a.raw_cap = unstaged ? 0 : static_cast<uint32_t>(((span_max * pb + 15) / 16) * 16 + 32);
    a.out = d_out;
    a.one = make_float2(1.0f, 1.0f);
    if (fuse) a.fft = *fuse;
    a.fuse_S = fuse ? fuse_stride : 0; // != 0: overlapping windows cut from this stream launch
    if (stage2) { // ... or from a second lowpass over it
        a.st2_L = stage2->L;
        a.st2_D = stage2->D;
        a.st2_total = stage2_total;
        for (uint32_t j = 0; j < stage2->L && j < 256; j++) a.st2_taps[j] = stage2->st->taps[j];
    }
    if (snap) {
        a.tail_out = snap->out;
        a.tail_W = snap->W, a.tail_S = snap->S, a.tail_T = snap->T;
        a.tail_units = snap->units;
    }
    {
        double rsum = 0.0;
        for (int i = 0; i < n_shift; i++) rsum += ratios[i];
        for (int k = 0; k < 4; k++) a.rot[k] = make_float2(static_cast<float>(cos(k * rsum)), static_cast<float>(sin(k * rsum)));
        const double step = 4.0 * lp.shape.NT * rsum;
        a.rot_step = make_float2(static_cast<float>(cos(step)), static_cast<float>(sin(step)));
        a.rot2c = static_cast<float>(2.0 * cos(rsum));
        if (n_shift == 1 && ratios[0] != 0.0 && std::isnormal(ratios[0])) {
            int ex = 0;
            const double m = frexp(fabs(ratios[0]), &ex); // |ratio| = m * 2^ex, m in [0.5, 1)
            a.rmant = static_cast<uint64_t>(ldexp(m, 53));
            a.rexp = ex - 53;
            a.rsign = ratios[0] < 0.0 ? -1 : 1;
        }
    }
    FirTaps taps;
    memset(&taps, 0, sizeof taps);
    const bool exact = c.precision == QD_PRECISION_EXACT;
    // FAST mode leaves cs8 samples as integers in the kernel and carries the 1/127 of lib.rs:251 in the taps
    const float scale = (!exact && fmt == QD_FMT_CS8) ? 1.0f / 127.0f : 1.0f;
    for (uint32_t j = 0; j < a.L; j++) {
        taps.t[j] = make_float2(st.taps[j] * scale, st.taps[j] * scale);
        taps.s[j] = st.taps[j] * scale;
    }
    switch (D) {
    case 2: return launch_fir_d2(c, a, taps, exact);
    case 4: return launch_fir_d4(c, a, taps, exact);
    case 8: return launch_fir_d8(c, a, taps, exact);
    case 16: return launch_fir_d16(c, a, taps, exact);
    case 32: return launch_fir_d32(c, a, taps, exact);
    }
    return set_error(QD_E_INVALID_ARG, "internal: no fused FIR for decimate %u", D);
}

// The truncated tails of nu windows (off0 + u*S, n_call) in the exact arithmetic, each at out[u*pitch + off + r]
static int launch_tail(Chain &c, const LpInfo &top, int fmt, int n_shift, const double *ratios, const uint8_t *d_src, uint64_t src_base,
                       uint64_t off0, uint64_t S, uint64_t n_call, uint64_t nu, float2 *out, uint64_t pitch, uint64_t off)
{
    TailArgs ta;
    memset(&ta, 0, sizeof ta);
    ta.src = d_src;
    ta.src_base = src_base;
    ta.fmt = fmt;
    ta.n_shift = n_shift;
    for (int i = 0; i < n_shift; i++) ta.ratio[i] = ratios[i];
    ta.sincos = c.ctx->d_sincos;
    ta.L = top.L, ta.D = top.D, ta.T = top.T;
    ta.span = top.T * top.D + top.L / 2;
    ta.log_d = 0;
    while ((1u << ta.log_d) < top.D) ta.log_d++;
    ta.off0 = off0, ta.S = S, ta.n_call = n_call, ta.n_units = nu;
    ta.out = out;
    ta.out_pitch = pitch, ta.out_off = off;
    ta.one = make_float2(1.0f, 1.0f);
    FirTaps tt;
    memset(&tt, 0, sizeof tt);
    for (uint32_t i = 0; i < top.L; i++) tt.t[i] = make_float2(top.st->taps[i], top.st->taps[i]);
    ta.wpw = std::max<uint32_t>(1, std::min<uint32_t>(32 / top.T, 8));
    const size_t tsm = static_cast<size_t>(kTailWarps) * ta.wpw * (ta.span + (ta.span >> ta.log_d) + 1) * sizeof(float2);
    if (tsm > 48 * 1024) QD_CUDA(cudaFuncSetAttribute(fk_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(tsm)));
    const uint64_t per_cta = static_cast<uint64_t>(kTailWarps) * ta.wpw;
    fk_tail<<<static_cast<unsigned>((nu + per_cta - 1) / per_cta), 32 * kTailWarps, tsm, c.stream>>>(ta, tt);
    QD_LAUNCHED();
    return QD_OK;
}

// FAST arithmetic over a cs8 capture, units that tile the output stream: the filter runs on the tensor cores
// (fk_tcfir) as one untruncated stream, and the T truncated outputs at the end of every unit are then written over
// by fk_tail.  *done = false: the shape is not one the tensor-core kernel takes.
static int run_tc(Chain &c, const FastPlan &f, const double *ratios, const uint8_t *d_src, uint64_t src_base, uint64_t src_end,
                  uint64_t soff, uint64_t unit_len, uint64_t stride, uint64_t nu, float2 *d_out, bool *done)
{
    *done = false;
    const LpInfo &top = f.lp[0];
    if (!c.use_tc || c.precision != QD_PRECISION_FAST || c.src.format != QD_FMT_CS8 || f.n_lp != 1 || f.stream_top) return QD_OK;
    if (nu > 1 && stride != unit_len) return QD_OK;
    if (top.T > 32 || top.T >= unit_len) return QD_OK;
    // measured (scripts/tc_shapes.py, 2^28 samples): the tensor-core kernel is 1.4 - 5.2 x the CUDA-core one for every
    // shape it takes except the 40-tap filter at decimate 4 (0.78 x: 16 outputs start in every row, and the CUDA-core
    // kernel has that filter unrolled at compile time)
    if (top.D <= 4 && top.L == 40) return QD_OK;
    TcGeom g;
    if (!tcfir_geometry(top.L, top.D, &g)) return QD_OK;
    double rsum = 0.0, rabs = 0.0;
    for (int i = 0; i < f.n_shift; i++) rsum += ratios[i], rabs += fabs(ratios[i]);
    // The reference mixes with cos/sin of fl64(n * ratio) (shift.rs:49), which is off the exact product by up to half
    // an ulp of n * ratio; the CUDA-core FAST kernel re-applies that per-sample rounding, a matrix product cannot.  It
    // is taken only while that phase error stays below 2^-25 rad (3e-8: n * ratio < 2^29, e.g. config 2's 2^30 samples),
    // where it is two orders of magnitude below the 1e-5 tolerance even on an output 60 dB under the input.
    {
        int ex = 0;
        frexp(static_cast<double>(src_end) * rabs, &ex);
        if (rabs != 0.0 && ex > 29) return QD_OK;
    }
    uint64_t key[4] = {top.L, top.D, 0, 0};
    memcpy(&key[2], &rsum, 8);
    for (uint32_t j = 0; j < top.L; j++) {
        uint32_t b;
        memcpy(&b, &top.st->taps[j], 4);
        key[3] = key[3] * 0x9E3779B97F4A7C15ull + b;
    }
    if (!c.tc_bimg.p || memcmp(key, c.tc_key, sizeof key) != 0) {
        QD_CUDA(cudaStreamSynchronize(c.stream)); // an earlier launch may still read the old image
        tcfir_b_image(g, top.st->taps.data(), top.L, top.D, rsum, c.tc_host, &c.tc_s_hi, &c.tc_s_lo);
        QD_TRY(c.ensure(c.tc_bimg, c.tc_host.size()));
        QD_CUDA(cudaMemcpyAsync(c.tc_bimg.p, c.tc_host.data(), c.tc_host.size(), cudaMemcpyHostToDevice, c.stream));
        QD_CUDA(cudaStreamSynchronize(c.stream));
        memcpy(c.tc_key, key, sizeof key);
    }
    QD_TRY(launch_tcfir(c, g, static_cast<const uint8_t *>(c.tc_bimg.p), c.tc_s_hi, c.tc_s_lo, top.L, top.D, f.n_shift, ratios, d_src,
                        src_base, src_end, soff, soff + nu * unit_len, d_out));
    if (top.T > 0)
        QD_TRY(launch_tail(c, top, c.src.format, f.n_shift, ratios, d_src, src_base, soff, unit_len, unit_len, nu, d_out, unit_len,
                           unit_len - top.T));
    *done = true;
    return QD_OK;
}

// Number of leading units of the arithmetic progression off0 + u*stride that are FULL: the read
// returns unit_len samples and its raw span lies inside the capture (so only the per-unit
// truncation rule applies, never the end-of-file one).
static uint64_t full_prefix(const Chain &c, uint64_t off0, uint64_t stride, uint64_t n_units, uint64_t unit_len)
{
    uint64_t need = unit_len; // raw samples an un-clamped read touches
    for (size_t s = c.stages.size(); s-- > 0;)
        if (c.stages[s].kind == QD_STAGE_LOWPASS) need = need * c.stages[s].decimate + c.stages[s].size;
    auto full = [&](uint64_t u) {
        uint64_t lo, hi, v = 0;
        const uint64_t off = off0 + u * stride;
        chain_source_span(c, off, unit_len, &lo, &hi);
        if (chain_valid(c, off, unit_len, &v) != QD_OK || v != unit_len) return false;
        return hi - lo == need;
    };
    if (n_units == 0) return 0;
    if (full(n_units - 1)) return n_units;
    if (!full(0)) return 0;
    uint64_t lo = 0, hi = n_units - 1; // lo full, hi not
    while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (full(mid)) lo = mid;
        else hi = mid;
    }
    return lo + 1;
}

static uint64_t round_up(uint64_t v, uint64_t m) { return (v + m - 1) / m * m; }

constexpr uint64_t kStreamCall = uint64_t(1) << 40; // "one read that never ends": no truncation inside a stream

// see qd_internal.h
int run_units_fast(Chain &c, uint64_t off0, uint64_t stride, uint64_t n_units, uint64_t unit_len, float2 *d_direct, FastSegmentFn on_segment, void *user, uint64_t *units_done, FastPrepareFn prepare)
{
    *units_done = 0;
    const FastPlan f = fast_plan(c, unit_len, stride, n_units);
    if (!f.ok) return QD_OK;
    const uint64_t n_full = full_prefix(c, off0, stride, n_units, unit_len);
    if (n_full == 0) return QD_OK;
    QD_CUDA(cudaSetDevice(c.device));
    const Source &s = c.src;
    const uint64_t pb = pair_bytes(s.format);
    const bool on_device = s.kind == QD_SRC_DEVICE_MEM;
    const LpInfo &top = f.lp[f.n_lp - 1];
    uint64_t mult = 1; // raw samples per top-level sample
    for (int k = 0; k < f.n_lp; k++) mult *= f.lp[k].D;
    if (f.stream_top) d_direct = nullptr; // streams are padded: never written into the caller's buffer
    double ratios[kMaxLeadShifts];
    for (int i = 0; i < f.n_shift; i++) ratios[i] = c.stages[i].ratio;

    // segment size: everything at once when both ends are resident on the device, else bounded
    // staging buffers that are double buffered against the copies
    uint64_t seg_units = n_full;
    // a segment of nu units stages the raw range of (nu - 1) * stride + unit_len top-level samples, and a stream
    // of the same length: both grow by `stride` per unit, also when the windows are further apart than wide
    const uint64_t step = (stride && n_units > 1) ? stride : unit_len;
    const uint64_t raw_per_unit = std::max<uint64_t>(1, step * mult * pb);
    if (!on_device) seg_units = std::max<uint64_t>(1, c.segment_bytes / raw_per_unit);
    if (!d_direct) {
        const uint64_t per_unit_out = (f.stream_top ? step : unit_len) * sizeof(float2) *
                                      (f.n_lp == 2 ? (1 + top.D) : 1);
        seg_units = std::min<uint64_t>(seg_units, std::max<uint64_t>(1, c.scratch_budget / 2 / std::max<uint64_t>(1, per_unit_out)));
    }
    seg_units = std::min(seg_units, n_full);

    QD_TRY(c.ensure_pipeline());
    // work queued earlier on the caller's stream precedes our copies
    QD_CUDA(cudaEventRecord(c.ev_entry, c.stream));
    QD_CUDA(cudaStreamWaitEvent(c.h2d_stream, c.ev_entry, 0));
    QD_CUDA(cudaStreamWaitEvent(c.d2h_stream, c.ev_entry, 0));

    uint64_t seg = 0;
    for (uint64_t u0 = 0; u0 < n_full; u0 += seg_units, ++seg) {
        const int j = static_cast<int>(seg & 1);
        const uint64_t nu = std::min(seg_units, n_full - u0);
        const uint64_t soff = off0 + u0 * stride;
        uint64_t lo, hi, lo2, hi2;
        chain_source_span(c, soff, unit_len, &lo, &hi);
        chain_source_span(c, soff + (nu - 1) * stride, unit_len, &lo2, &hi2);
        hi = std::max(hi, hi2);
        if (lo < s.base_sample || hi > s.base_sample + s.resident_samples)
            return set_error(QD_E_NOT_RESIDENT, "samples [%llu, %llu) requested but this source holds [%llu, %llu)",
                             (unsigned long long)lo, (unsigned long long)hi, (unsigned long long)s.base_sample,
                             (unsigned long long)(s.base_sample + s.resident_samples));
        const uint8_t *d_src;
        uint64_t src_base, src_end;
        if (on_device) {
            d_src = s.data;
            src_base = s.base_sample;
            src_end = s.base_sample + s.resident_samples;
        } else {
            // stage [lo, hi) at the byte offset that keeps absolute sample 0 on a 16-byte boundary
            const size_t skew = static_cast<size_t>((lo * pb) & 15);
            const size_t bytes = static_cast<size_t>((hi - lo) * pb);
            QD_TRY(c.ensure(c.pipe_in[j], bytes + 64));
            uint8_t *dst = static_cast<uint8_t *>(c.pipe_in[j].p) + skew;
            if (seg >= 2) QD_CUDA(cudaStreamWaitEvent(c.h2d_stream, c.ev_compute[j], 0)); // buffer j is free again
            if (s.kind == QD_SRC_HOST_MEM) {
                QD_CUDA(cudaMemcpyAsync(dst, s.data + (lo - s.base_sample) * pb, bytes, cudaMemcpyHostToDevice, c.h2d_stream));
            } else {
                QD_TRY(c.ensure_pinned2(j, bytes));
                if (seg >= 2) QD_CUDA(cudaEventSynchronize(c.ev_h2d[j])); // pinned buffer j has been consumed
                size_t done = 0;
                while (done < bytes) {
                    const ssize_t r = pread(s.fd, static_cast<uint8_t *>(c.h_pin2[j]) + done, bytes - done,
                                            static_cast<off_t>(lo * pb + done));
                    if (r <= 0) return set_error(QD_E_IO, "read %s: %s", s.path.c_str(), r < 0 ? strerror(errno) : "unexpected end of file");
                    done += static_cast<size_t>(r);
                }
                QD_CUDA(cudaMemcpyAsync(dst, c.h_pin2[j], bytes, cudaMemcpyHostToDevice, c.h2d_stream));
            }
            QD_CUDA(cudaEventRecord(c.ev_h2d[j], c.h2d_stream));
            QD_CUDA(cudaStreamWaitEvent(c.stream, c.ev_h2d[j], 0));
            d_src = static_cast<const uint8_t *>(c.pipe_in[j].p) + ((lo * pb) & 15);
            src_base = lo;
            src_end = hi;
        }
        if (seg >= 2) QD_CUDA(cudaStreamWaitEvent(c.stream, c.ev_d2h[j], 0)); // output staging j drained

        const float2 *d_top;
        uint64_t pitch;
        bool tc_done = false; // the segment's filter ran on the tensor cores (fk_tcfir)
        QD_TRY(c.prof_begin());
        // sparkfft over back-to-back windows behind a run-time-length filter in EXACT arithmetic: the STFT runs inside
        // the filter kernel (whole windows per tile, window starts on multiples of R), nothing but glyph rows is written
        const uint64_t t_tile = static_cast<uint64_t>(top.shape.R) * top.shape.NT;
        const bool fuse = c.fuse_stft != 0 && !f.stream_top && prepare && f.n_lp == 1 && stride == unit_len && is_pow2(unit_len) && unit_len >= 4 &&
                          unit_len <= t_tile && top.L != 40 && c.precision == QD_PRECISION_EXACT && unit_len % top.shape.R == 0 &&
                          soff % unit_len == 0;
        if (fuse) {
            FftArgs fa;
            QD_TRY(prepare(c, user, j, u0, nu, &fa));
            QD_TRY(launch_fir(c, top, s.format, f.n_shift, ratios, d_src, src_base, src_end, soff, unit_len, stride, nu,
                              nu * unit_len, nullptr, nullptr, &fa));
            d_top = nullptr;
            pitch = unit_len;
        } else if (!f.stream_top) {
            // [nu][unit_len] matrix straight from the raw bytes, per-unit truncation applied in the kernel
            float2 *d_out;
            if (d_direct) {
                d_out = d_direct + u0 * unit_len;
            } else {
                QD_TRY(c.ensure(c.pipe_out[j], nu * unit_len * sizeof(float2)));
                d_out = static_cast<float2 *>(c.pipe_out[j].p);
            }
            QD_TRY(run_tc(c, f, ratios, d_src, src_base, src_end, soff, unit_len, stride, nu, d_out, &tc_done));
            if (!tc_done)
                QD_TRY(launch_fir(c, top, s.format, f.n_shift, ratios, d_src, src_base, src_end, soff, unit_len, stride, nu,
                                  nu * unit_len, d_out));
            d_top = d_out;
            pitch = unit_len;
        } else {
            // top-level outputs [g0, g1) as one stream; windows are cut from it at the sink's stride
            const uint64_t g0 = soff, g1 = soff + (nu - 1) * stride + unit_len;
            const uint64_t glen = round_up(g1 - g0, top.shape.R);
            // Truncated tails: when every stream output belongs to at most one window's tail (stride >= T) they are
            // snapshots of the stream kernel's own running sums; otherwise fk_tail computes them on their own.
            const bool snap_tails = f.stream_tail && stride >= top.T;
            // sparkfft inside the (top) stream kernel: windows carried from tile to tile (fk_fir FUSE = 2)
            const uint64_t tt = tile_outputs(static_cast<int>(top.D), top.shape.R, top.shape.NT, top.L == 40 ? 40 : 0);
            // Opt-in ("fuse_stft" = 2), because it measured SLOWER than the separate kernels on every benchmark shape:
            // the per-tile epilogue (carry exchange, cooperative radix-4 passes over shared memory, barriers) is work
            // the dedicated fk_stft kernel does better with register-blocked passes, and in the short-filter stream
            // kernels -- which run at the HBM rate -- it stalls the tile loop that keeps loads in flight.  Same box,
            // Gsamples/s: config 2's input through 64/16 windows 179 -> 111, config 1 130 -> 114, config 5 542 -> 517
            // (STFT in the second filter's kernel) or 345 (both filters and the STFT in one kernel).
            const bool fuse_stream = c.fuse_stft == 2 && prepare &&
                                     c.precision == QD_PRECISION_EXACT && is_pow2(unit_len) && unit_len >= 4 &&
                                     unit_len - 1 <= tt && stride >= 1 && (top.T == 0 || snap_tails) &&
                                     fuse_scratch_fits(top, unit_len, stride) && glen * (f.n_lp == 2 ? top.D : 1) + top.L < (uint64_t(1) << 31) &&
                                     nu < (uint64_t(1) << 31);
            if (fuse_stream) {
                FftArgs fa;
                QD_TRY(prepare(c, user, j, u0, nu, &fa));
                TailSnap ts{nullptr, static_cast<uint32_t>(unit_len), static_cast<uint32_t>(stride), top.T, nu};
                if (f.n_lp == 1) {
                    QD_TRY(launch_fir(c, top, s.format, f.n_shift, ratios, d_src, src_base, src_end, g0, kStreamCall, kStreamCall,
                                      1, glen, nullptr, &ts, &fa, static_cast<uint32_t>(stride)));
                } else if (fuse_two_stage_fits(f.lp[0], top, unit_len, stride)) {
                    // both filters and the STFT in ONE kernel: the inner stage's stream kernel carries its outputs, the
                    // outer filter and the windows run on what each tile completes (FirArgs::st2_L)
                    const LpInfo &in = f.lp[0];
                    const uint64_t h0 = g0 * top.D + top.i0, h1 = (g1 - 1) * top.D + top.i0 + top.L;
                    const uint64_t hlen = round_up(h1 - h0, in.shape.R);
                    ts.T = 0;
                    QD_TRY(launch_fir(c, in, s.format, f.n_shift, ratios, d_src, src_base, src_end, h0, kStreamCall, kStreamCall, 1,
                                      hlen, nullptr, &ts, &fa, static_cast<uint32_t>(stride), &top, g1 - g0));
                } else {
                    const LpInfo &in = f.lp[0];
                    const uint64_t h0 = g0 * top.D + top.i0, h1 = (g1 - 1) * top.D + top.i0 + top.L;
                    const uint64_t hlen = round_up(h1 - h0, in.shape.R);
                    QD_TRY(c.ensure(c.pipe_mid[j], hlen * sizeof(float2) + 64));
                    float2 *d_mid = reinterpret_cast<float2 *>(static_cast<uint8_t *>(c.pipe_mid[j].p) + ((h0 & 1) ? 8 : 0));
                    QD_TRY(launch_fir(c, in, s.format, f.n_shift, ratios, d_src, src_base, src_end, h0, kStreamCall, kStreamCall, 1,
                                      hlen, d_mid));
                    QD_TRY(launch_fir(c, top, QD_FMT_CF32, 0, nullptr, reinterpret_cast<const uint8_t *>(d_mid), h0, h1, g0,
                                      kStreamCall, kStreamCall, 1, glen, nullptr, &ts, &fa, static_cast<uint32_t>(stride)));
                }
                QD_TRY(c.prof_end("fk_fir (fused decode+mix+FIR-decimate+STFT+glyphs)"));
                QD_CUDA(cudaEventRecord(c.ev_compute[j], c.stream));
                if (on_segment) QD_TRY(on_segment(c, user, j, u0, nu, nullptr, stride));
                continue;
            }
            QD_TRY(c.ensure(c.pipe_out[j], glen * sizeof(float2) + 64));
            float2 *d_out = static_cast<float2 *>(c.pipe_out[j].p);
            if (f.stream_tail) {
                QD_TRY(c.ensure(c.pipe_tail[j], nu * top.T * sizeof(float2)));
                c.seg_tail = static_cast<float2 *>(c.pipe_tail[j].p);
                c.seg_tail_len = top.T;
            }
            if (f.n_lp == 1) {
                TailSnap ts{static_cast<float2 *>(c.pipe_tail[j].p), static_cast<uint32_t>(unit_len), static_cast<uint32_t>(stride),
                            top.T, nu};
                QD_TRY(launch_fir(c, top, s.format, f.n_shift, ratios, d_src, src_base, src_end, g0, kStreamCall, kStreamCall,
                                  1, glen, d_out, snap_tails ? &ts : nullptr));
            } else {
                const LpInfo &in = f.lp[0];
                // inner outputs the outer stage reads for [g0, g1): h = g*D2 + i0_2 + j
                const uint64_t h0 = g0 * top.D + top.i0, h1 = (g1 - 1) * top.D + top.i0 + top.L;
                const uint64_t hlen = round_up(h1 - h0, in.shape.R);
                QD_TRY(c.ensure(c.pipe_mid[j], hlen * sizeof(float2) + 64));
                // keep absolute inner index 0 on a 16-byte boundary: the stream starts 8 bytes in when h0 is odd
                float2 *d_mid = reinterpret_cast<float2 *>(static_cast<uint8_t *>(c.pipe_mid[j].p) + ((h0 & 1) ? 8 : 0));
                QD_TRY(launch_fir(c, in, s.format, f.n_shift, ratios, d_src, src_base, src_end, h0, kStreamCall, kStreamCall, 1,
                                  hlen, d_mid));
                QD_TRY(launch_fir(c, top, QD_FMT_CF32, 0, nullptr, reinterpret_cast<const uint8_t *>(d_mid), h0, h1, g0,
                                  kStreamCall, kStreamCall, 1, glen, d_out));
            }
            d_top = d_out;
            pitch = stride;
            if (f.stream_tail && !snap_tails)
                QD_TRY(launch_tail(c, top, s.format, f.n_shift, ratios, d_src, src_base, soff, stride, unit_len, nu,
                                   static_cast<float2 *>(c.pipe_tail[j].p), top.T, 0));
        }
        QD_TRY(c.prof_end(fuse      ? "fk_fir (fused decode+mix+FIR-decimate+STFT+glyphs)"
                          : tc_done ? "fk_tcfir (tensor-core decode+mix+FIR-decimate) + fk_tail"
                                    : "fk_fir (fused decode+mix+FIR-decimate)"));
        QD_CUDA(cudaEventRecord(c.ev_compute[j], c.stream)); // the raw staging buffer may be refilled
        if (on_segment) {
            const int rc = on_segment(c, user, j, u0, nu, d_top, pitch);
            c.seg_tail = nullptr;
            c.seg_tail_len = 0;
            if (rc != QD_OK) return rc;
        }
    }
    *units_done = n_full;
    return QD_OK;
}

} // namespace qd
