// qd_generic.cu -- the general ("unit-local") executor: any chain, any order of stages.
//
// A sink unit is one top-level Samples::read_at(off, n) of the reference: a sparkfft window
// (fft.rs:29-30), a do_write chunk (lib.rs:201-202), a take_fft row (ffts.rs:62) or a caller's own
// read_at.  Each unit is evaluated exactly as the reference evaluates it -- every stage sees the
// (off, n) its outer stage would have issued (filter.rs:68-71), including the zero-truncated filter
// tail at the end of each LowPass::read_at buffer (filter.rs:107-124) -- but all units of a batch
// run in parallel and every stage is a kernel over [units x samples].  This path is the semantic
// ground truth on the GPU; qd_fast.cu holds the fused kernels for the canonical chains.
#include <algorithm>
#include <cstring>

#include <unistd.h>

#include "qd_device_math.cuh"
#include "qd_internal.h"
#include "qd_stft_epilogue.cuh"

namespace qd {

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t unit_off_top(uint64_t off0, uint64_t stride, const uint64_t *__restrict__ offsets,
                                                 uint64_t u)
{
    return offsets ? offsets[u] : off0 + u * stride;
}

// valid[level * B + u] = samples the reference's read_at returns at that level for unit u
__global__ void gk_geometry(GPlan p, uint64_t off0, uint64_t stride, const uint64_t *__restrict__ offsets, uint32_t B,
                            uint32_t *__restrict__ valid)
{
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= B) return;
    const uint64_t o = unit_off_top(off0, stride, offsets, u) * p.mult[0];
    uint64_t v;
    if (p.src_kind == QD_SRC_GEN) v = p.n_level[0];            // gen.rs:36,46: always fills the buffer
    else v = o < p.src_total ? min(p.n_level[0], p.src_total - o) : 0; // samples.rs:80-93
    valid[u] = static_cast<uint32_t>(v);
    for (int s = 0; s < p.n_stages; s++) {
        if (p.st[s].kind == QD_STAGE_LOWPASS) v = v >= p.st[s].L ? (v - p.st[s].L) / p.st[s].D : 0; // filter.rs:76
        valid[static_cast<size_t>(s + 1) * B + u] = static_cast<uint32_t>(v);
    }
}

// Gen::read_at, gen.rs:36-44
__device__ __forceinline__ float2 gen_sample(const GPlan &p, uint64_t n)
{
    const double tau = 6.283185307179586; // PI * 2.
    const double base = __ddiv_rn(__dmul_rn(__ull2double_rn(n), tau), __ull2double_rn(p.src_rate));
    float2 val = make_float2(0.0f, 0.0f);
    for (int t = 0; t < p.n_tones; t++) {
        const double f = __dmul_rn(__ll2double_rn(p.tones[t]), base);
        double c, s;
        sincos_f64(f, p.sincos, c, s);
        val.x = __fadd_rn(val.x, static_cast<float>(c));
        val.y = __fadd_rn(val.y, static_cast<float>(s));
    }
    return val;
}

// level 0: SampleFile::read_at / Gen::read_at, with the shifts that directly follow the source
// applied in the same pass (Shift::read_at, shift.rs:46-54)
__global__ void gk_source(GPlan p, int n_lead_shifts, uint64_t off0, uint64_t stride,
                          const uint64_t *__restrict__ offsets, uint32_t B, const uint32_t *__restrict__ valid,
                          float2 *__restrict__ out, uint32_t blocks_per_unit)
{
    const uint32_t u = blockIdx.x / blocks_per_unit;
    const uint64_t i = static_cast<uint64_t>(blockIdx.x % blocks_per_unit) * blockDim.x + threadIdx.x;
    const uint64_t n0 = p.n_level[0];
    if (u >= B || i >= n0) return;
    float2 v = make_float2(0.0f, 0.0f); // the reference's buffers start zeroed
    if (i < valid[u]) {
        const uint64_t n = unit_off_top(off0, stride, offsets, u) * p.mult[0] + i;
        v = p.src_kind == QD_SRC_GEN ? gen_sample(p, n) : decode_sample(p.src, p.fmt, n - p.src_base);
        for (int s = 0; s < n_lead_shifts; s++) v = cmul_exact(v, phasor_exact(n, p.st[s].ratio, p.sincos));
    }
    out[static_cast<size_t>(u) * n0 + i] = v;
}

// a Shift that follows a LowPass: in place on its level
__global__ void gk_shift(GPlan p, int stage, uint64_t off0, uint64_t stride, const uint64_t *__restrict__ offsets,
                         uint32_t B, const uint32_t *__restrict__ valid, float2 *__restrict__ buf,
                         uint32_t blocks_per_unit)
{
    const int level = stage + 1;
    const uint32_t u = blockIdx.x / blocks_per_unit;
    const uint64_t i = static_cast<uint64_t>(blockIdx.x % blocks_per_unit) * blockDim.x + threadIdx.x;
    if (u >= B || i >= valid[static_cast<size_t>(level) * B + u]) return;
    const uint64_t n = unit_off_top(off0, stride, offsets, u) * p.mult[level] + i;
    float2 *q = buf + static_cast<size_t>(u) * p.n_level[level] + i;
    *q = cmul_exact(*q, phasor_exact(n, p.st[stage].ratio, p.sincos));
}

// LowPass::read_at, filter.rs:54-83, kept outputs only:
//   y[k] = sum_{j < min(L, valid_in - k*D - i0)} raw[k*D + i0 + j] * f[j],  i0 = L - L/2,
// ascending j, multiply then add, each rounded (filter.rs:112-120; convoluted[L + k*D], :78-80).
__global__ void gk_lowpass(GPlan p, int stage, uint32_t B, const uint32_t *__restrict__ valid,
                           const float2 *__restrict__ in, float2 *__restrict__ out, uint32_t blocks_per_unit)
{
    const uint32_t u = blockIdx.x / blocks_per_unit;
    const uint64_t k = static_cast<uint64_t>(blockIdx.x % blocks_per_unit) * blockDim.x + threadIdx.x;
    const uint64_t n_in = p.n_level[stage], n_out = p.n_level[stage + 1];
    if (u >= B || k >= n_out) return;
    float2 acc = make_float2(0.0f, 0.0f);
    if (k < valid[static_cast<size_t>(stage + 1) * B + u]) {
        const uint32_t L = p.st[stage].L;
        const uint64_t base = k * p.st[stage].D + (L - L / 2);
        const uint64_t v_in = valid[static_cast<size_t>(stage) * B + u];
        const uint32_t J = static_cast<uint32_t>(min(static_cast<uint64_t>(L), v_in - base));
        const float2 *x = in + static_cast<size_t>(u) * n_in + base;
        const float *__restrict__ f = p.st[stage].taps;
        for (uint32_t j = 0; j < J; j++) {
            const float2 s = x[j];
            const float t = __ldg(f + j);
            acc.x = __fadd_rn(acc.x, __fmul_rn(s.x, t));
            acc.y = __fadd_rn(acc.y, __fmul_rn(s.y, t));
        }
    }
    out[static_cast<size_t>(u) * n_out + k] = acc;
}

// ---- FFT of one unit per CTA in shared memory: our radix-4 DIT definition -------------------
// A team of `team` threads transforms one window in shared memory; a CTA holds blockDim/team windows.
__global__ void gk_fft(FftArgs a)
{
    extern __shared__ float2 fft_smem[];
    const uint32_t W = a.W, TW = a.team;
    const uint32_t wpc = blockDim.x / TW;
    const uint32_t team = threadIdx.x / TW, lt = threadIdx.x - team * TW;
    const uint64_t u = static_cast<uint64_t>(blockIdx.x) * wpc + team;
    const bool active = u < a.n_units;
    float2 *x = fft_smem + static_cast<size_t>(team) * W;
    const int logw = 31 - __clz(W);
    const bool odd = logw & 1;
    const int n_r4 = logw >> 1;
    if (active) {
        for (uint32_t n = lt; n < W; n += TW) {
            float2 v;
            if (a.raw) v = decode_sample(a.raw, a.raw_fmt, a.raw_first + u * a.in_pitch + n);
            else v = a.in[u * a.in_pitch + n];
            if (a.window) { // ffts.rs:64-68: Complex<f32> *= f32
                const float w = a.window[n];
                v = make_float2(__fmul_rn(v.x, w), __fmul_rn(v.y, w));
            }
            x[leaf_position(n, W, n_r4, odd)] = v;
        }
    }
    __syncthreads();
    if (odd) { // innermost size-2 FFTs
        if (active)
            for (uint32_t b = lt; b < W / 2; b += TW) {
                const float2 p = x[2 * b], q = x[2 * b + 1];
                x[2 * b] = cadd(p, q);
                x[2 * b + 1] = csub(p, q);
            }
        __syncthreads();
    }
    for (uint32_t q = odd ? 2 : 1; q < W; q <<= 2) {
        const uint32_t scale = W / (4 * q); // w(4q, j) == w(W, j * W/(4q))
        if (active)
            for (uint32_t b = lt; b < W / 4; b += TW) {
                const uint32_t blk = b / q, k = b - blk * q;
                float2 *base = x + static_cast<size_t>(blk) * 4 * q + k;
                float2 t0 = base[0], t1 = base[q], t2 = base[2 * q], t3 = base[3 * q];
                if (k != 0) {
                    t1 = cmul_tw(t1, __ldg(a.tw + k * scale));
                    t2 = cmul_tw(t2, __ldg(a.tw + 2 * k * scale));
                    t3 = cmul_tw(t3, __ldg(a.tw + 3 * k * scale));
                }
                radix4(t0, t1, t2, t3);
                base[0] = t0;
                base[q] = t1;
                base[2 * q] = t2;
                base[3 * q] = t3;
            }
        __syncthreads();
    }
    if (!active) return;
    if (a.epi == EPI_LEVELS) { // fft.rs:95-97: sequential f32 sums over the natural-order halves
        if (lt == 0) {
            float first = 0.0f, second = 0.0f;
            for (uint32_t b = 0; b < W / 2; b++) first = __fadd_rn(first, hypot_exact(x[b].x, x[b].y));
            for (uint32_t b = W / 2; b < W; b++) second = __fadd_rn(second, hypot_exact(x[b].x, x[b].y));
            a.idx[u] = first < second ? 0 : 1;
        }
        return;
    }
    const uint32_t half = W / 2;
    for (uint32_t b = lt; b < W; b += TW) {
        // iter().skip(w/2).chain(iter().take(w/2)), fft.rs:48-52 / ffts.rs:72-76
        const uint32_t src = b < W - half ? b + half : b - (W - half);
        const float norm = hypot_exact(x[src].x, x[src].y);
        const size_t o = static_cast<size_t>(u) * W + b;
        if (a.mag) a.mag[o] = norm;
        if (a.epi == EPI_SPARK) {
            const int g = glyph_index(norm, a.mn, a.mx, a.distinction);
            if (g == 9) *a.panic_flag = 1;
            a.idx[o] = static_cast<uint8_t>(g);
        }
    }
}

// ---- any width that is not a power of two (take_fft only: FftPlanner accepts any length, ffts.rs:25) ----
// Direct DFT  X[k] = sum_j x[j] * w(W, (j*k) mod W), products and running sum in f64 (each f32 x f32 product
// is exact in f64, so every term costs one rounding), ascending j, rounded to f32 at the end -- the same
// definition as oracle/quadrs_oracle.c plan_process_any, hence bit-identical.  One CTA per window.
__global__ void gk_dft(FftArgs a)
{
    extern __shared__ float2 dft_smem[];
    const uint32_t W = a.W;
    const uint64_t u = blockIdx.x;
    for (uint32_t n = threadIdx.x; n < W; n += blockDim.x) {
        float2 v = a.in[u * a.in_pitch + n];
        if (a.window) {
            const float w = a.window[n];
            v = make_float2(__fmul_rn(v.x, w), __fmul_rn(v.y, w));
        }
        dft_smem[n] = v;
    }
    __syncthreads();
    const uint32_t half = W / 2;
    for (uint32_t k = threadIdx.x; k < W; k += blockDim.x) {
        double re = 0.0, im = 0.0;
        uint32_t m = 0; // (j * k) mod W
        for (uint32_t j = 0; j < W; j++) {
            const float2 x = dft_smem[j];
            const float2 w = __ldg(a.tw + m);
            const double xr = x.x, xi = x.y, wr = w.x, wi = w.y;
            re = fma(xr, wr, re); // exact product, one rounding: re + xr*wr
            re = fma(-xi, wi, re);
            im = fma(xr, wi, im);
            im = fma(xi, wr, im);
            m += k;
            if (m >= W) m -= W;
        }
        const float norm = hypot_exact(static_cast<float>(re), static_cast<float>(im));
        // iter().skip(w/2).chain(iter().take(w/2)), ffts.rs:72-76: bin k lands at display column b
        const uint32_t b = k >= half ? k - half : k + (W - half);
        a.mag[static_cast<size_t>(u) * W + b] = norm;
    }
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------

Chain::~Chain()
{
    for (qd_chain *s : shards) delete s; // a sharded handle owns one chain per device
    shards.clear();
    if (src.fd >= 0) close(src.fd);
    src.fd = -1;
    if (!ctx) return; // description-only chain (qd_shard_plan): nothing lives on a device
    cudaSetDevice(device);
    auto rel = [](Buf &b) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    };
    rel(stage_in);
    for (auto &l : level) rel(l);
    rel(geo);
    rel(sink_a);
    rel(sink_b);
    rel(offsets);
    rel(twiddles);
    rel(window);
    rel(flag);
    rel(tc_bimg);
    for (int j = 0; j < 2; j++) {
        rel(pipe_in[j]);
        rel(pipe_mid[j]);
        rel(pipe_out[j]);
        rel(pipe_idx[j]);
        rel(pipe_mag[j]);
        rel(pipe_tail[j]);
        if (h_pin2[j]) cudaFreeHost(h_pin2[j]);
    }
    if (pipeline_ready) {
        cudaStreamDestroy(h2d_stream);
        cudaStreamDestroy(d2h_stream);
        cudaEvent_t evs[] = {ev_entry, ev_exit, ev_h2d[0], ev_h2d[1], ev_compute[0], ev_compute[1],
                             ev_sink[0], ev_sink[1], ev_d2h[0], ev_d2h[1]};
        for (cudaEvent_t e : evs) cudaEventDestroy(e);
    }
    for (auto &s : stages)
        if (s.d_taps) cudaFree(s.d_taps);
    if (h_pinned) cudaFreeHost(h_pinned);
    for (auto &e : prof_events) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (own_stream && stream) cudaStreamDestroy(stream);
}

int Chain::prof_begin()
{
    if (!profile) return QD_OK;
    if (prof_used == prof_events.size()) {
        cudaEvent_t a, b;
        QD_CUDA(cudaEventCreate(&a));
        QD_CUDA(cudaEventCreate(&b));
        prof_events.emplace_back(a, b);
    }
    QD_CUDA(cudaEventRecord(prof_events[prof_used].first, stream));
    return QD_OK;
}

int Chain::prof_end(const char *kernel)
{
    if (!profile) return QD_OK;
    QD_CUDA(cudaEventRecord(prof_events[prof_used].second, stream));
    if (prof_names.size() < prof_events.size()) prof_names.resize(prof_events.size());
    prof_names[prof_used] = kernel;
    prof_used++;
    return QD_OK;
}

int Chain::ensure(Buf &b, size_t bytes)
{
    if (bytes <= b.cap) return QD_OK;
    if (b.p) {
        QD_CUDA(cudaStreamSynchronize(stream));
        QD_CUDA(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    const size_t want = std::max(bytes, size_t(4096));
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return set_error(QD_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return QD_OK;
}

int Chain::ensure_pipeline()
{
    if (pipeline_ready) return QD_OK;
    QD_CUDA(cudaStreamCreateWithFlags(&h2d_stream, cudaStreamNonBlocking));
    QD_CUDA(cudaStreamCreateWithFlags(&d2h_stream, cudaStreamNonBlocking));
    cudaEvent_t *evs[] = {&ev_entry, &ev_exit, &ev_h2d[0], &ev_h2d[1], &ev_compute[0], &ev_compute[1],
                          &ev_sink[0], &ev_sink[1], &ev_d2h[0], &ev_d2h[1]};
    for (cudaEvent_t *e : evs) QD_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    pipeline_ready = true;
    return QD_OK;
}

int Chain::ensure_pinned2(int j, size_t bytes)
{
    if (bytes <= h_pin2_cap[j]) return QD_OK;
    if (h_pin2[j]) {
        QD_CUDA(cudaStreamSynchronize(h2d_stream));
        QD_CUDA(cudaFreeHost(h_pin2[j]));
        h_pin2[j] = nullptr;
        h_pin2_cap[j] = 0;
    }
    QD_CUDA(cudaMallocHost(&h_pin2[j], bytes));
    h_pin2_cap[j] = bytes;
    return QD_OK;
}

int Chain::ensure_pinned(size_t bytes)
{
    if (bytes <= h_pinned_cap) return QD_OK;
    if (h_pinned) {
        QD_CUDA(cudaStreamSynchronize(stream));
        QD_CUDA(cudaFreeHost(h_pinned));
        h_pinned = nullptr;
        h_pinned_cap = 0;
    }
    QD_CUDA(cudaMallocHost(&h_pinned, bytes));
    h_pinned_cap = bytes;
    return QD_OK;
}

// ---- the reference's count arithmetic, on the host ----

int chain_len(const Chain &c, uint64_t *len)
{
    uint64_t v = c.src.kind == QD_SRC_GEN
                     ? f64_as_u64(c.src.gen_seconds * static_cast<double>(c.src.sample_rate)) // gen.rs:31-33
                     : c.src.total_samples;                                                    // samples.rs:64-66
    for (const Stage &s : c.stages) {
        if (s.kind != QD_STAGE_LOWPASS) continue; // shift.rs:38-40
        if (!(v >= s.size))                       // filter.rs:46
            return set_error(QD_E_SHORT_INPUT, "assertion failed: self.inner.len() >= self.filter.len() as u64");
        v = 1 + (v - s.size) / s.decimate; // filter.rs:47
    }
    *len = v;
    return QD_OK;
}

uint64_t chain_rate(const Chain &c)
{
    uint64_t r = c.src.sample_rate;
    for (const Stage &s : c.stages)
        if (s.kind == QD_STAGE_LOWPASS) r = r / s.decimate; // filter.rs:50-52
    return r;
}

static void level_geometry(const Chain &c, uint64_t unit_len, uint64_t *n_level, uint64_t *mult)
{
    const int S = static_cast<int>(c.stages.size());
    n_level[S] = unit_len;
    mult[S] = 1;
    for (int s = S - 1; s >= 0; s--) {
        const Stage &st = c.stages[s];
        if (st.kind == QD_STAGE_LOWPASS) { // filter.rs:68-71
            n_level[s] = n_level[s + 1] * st.decimate + st.size;
            mult[s] = mult[s + 1] * st.decimate;
        } else {
            n_level[s] = n_level[s + 1];
            mult[s] = mult[s + 1];
        }
    }
}

int chain_valid(const Chain &c, uint64_t off, uint64_t n, uint64_t *valid)
{
    uint64_t n_level[kMaxStages + 1], mult[kMaxStages + 1];
    level_geometry(c, n, n_level, mult);
    uint64_t v;
    if (c.src.kind == QD_SRC_GEN) {
        v = n_level[0];
    } else {
        const uint64_t o = off * mult[0];
        if (!(o < c.src.total_samples)) // samples.rs:74
            return set_error(QD_E_OFFSET_EOF, "assertion failed: off < self.len() (off %llu, len %llu)",
                             (unsigned long long)o, (unsigned long long)c.src.total_samples);
        v = std::min(n_level[0], c.src.total_samples - o);
    }
    for (const Stage &s : c.stages) {
        if (s.kind != QD_STAGE_LOWPASS) continue;
        if (v < s.size) // filter.rs:76: usize underflow -> panic
            return set_error(QD_E_SHORT_INPUT, "attempt to subtract with overflow (valid %llu < filter %llu)",
                             (unsigned long long)v, (unsigned long long)s.size);
        v = (v - s.size) / s.decimate;
    }
    *valid = v;
    return QD_OK;
}

void chain_source_span(const Chain &c, uint64_t off, uint64_t n, uint64_t *lo, uint64_t *hi)
{
    uint64_t n_level[kMaxStages + 1], mult[kMaxStages + 1];
    level_geometry(c, n, n_level, mult);
    const uint64_t total = c.src.total_samples;
    *lo = std::min(off * mult[0], total);
    *hi = std::min(off * mult[0] + n_level[0], total);
}

static int fill_plan(Chain &c, uint64_t unit_len, GPlan *p)
{
    memset(p, 0, sizeof *p);
    const int S = static_cast<int>(c.stages.size());
    p->n_stages = S;
    p->src_kind = c.src.kind;
    p->fmt = c.src.format;
    level_geometry(c, unit_len, p->n_level, p->mult);
    for (int l = 0; l <= S; l++)
        if (p->n_level[l] >= (uint64_t(1) << 31))
            return set_error(QD_E_INVALID_ARG, "read of %llu samples needs %llu samples at level %d: too large",
                             (unsigned long long)unit_len, (unsigned long long)p->n_level[l], l);
    for (int s = 0; s < S; s++) {
        const Stage &st = c.stages[s];
        p->st[s].kind = st.kind;
        p->st[s].L = static_cast<uint32_t>(st.size);
        p->st[s].D = st.decimate;
        p->st[s].ratio = st.ratio;
        p->st[s].taps = st.d_taps;
    }
    p->src_total = c.src.total_samples;
    p->src_rate = c.src.sample_rate;
    p->n_tones = static_cast<int>(c.src.gen_cos.size());
    for (int t = 0; t < p->n_tones; t++) p->tones[t] = c.src.gen_cos[t];
    p->sincos = c.ctx->d_sincos;
    return QD_OK;
}

// Makes raw samples [lo, hi) available on the device; sets plan->src / src_base.
static int stage_source(Chain &c, uint64_t lo, uint64_t hi, GPlan *p)
{
    const Source &s = c.src;
    if (s.kind == QD_SRC_GEN || hi <= lo) {
        p->src = nullptr;
        p->src_base = 0;
        return QD_OK;
    }
    const uint64_t pb = pair_bytes(s.format);
    if (lo < s.base_sample || hi > s.base_sample + s.resident_samples)
        return set_error(QD_E_NOT_RESIDENT, "samples [%llu, %llu) requested but this source holds [%llu, %llu)",
                         (unsigned long long)lo, (unsigned long long)hi, (unsigned long long)s.base_sample,
                         (unsigned long long)(s.base_sample + s.resident_samples));
    if (s.kind == QD_SRC_DEVICE_MEM) {
        p->src = s.data;
        p->src_base = s.base_sample;
        return QD_OK;
    }
    const size_t bytes = static_cast<size_t>((hi - lo) * pb);
    QD_TRY(c.ensure(c.stage_in, bytes));
    if (s.kind == QD_SRC_HOST_MEM) {
        QD_CUDA(cudaMemcpyAsync(c.stage_in.p, s.data + (lo - s.base_sample) * pb, bytes, cudaMemcpyHostToDevice,
                                c.stream));
    } else { // FILE: one pread per staged range, as SampleFile::read_at does per call (samples.rs:80-83)
        QD_TRY(c.ensure_pinned(bytes));
        QD_CUDA(cudaStreamSynchronize(c.stream)); // the pinned buffer may still feed an earlier copy
        size_t done = 0;
        while (done < bytes) {
            const ssize_t r = pread(s.fd, static_cast<uint8_t *>(c.h_pinned) + done, bytes - done,
                                    static_cast<off_t>(lo * pb + done));
            if (r < 0) return set_error(QD_E_IO, "read %s: %s", s.path.c_str(), strerror(errno));
            if (r == 0) return set_error(QD_E_IO, "read %s: unexpected end of file", s.path.c_str());
            done += static_cast<size_t>(r);
        }
        QD_CUDA(cudaMemcpyAsync(c.stage_in.p, c.h_pinned, bytes, cudaMemcpyHostToDevice, c.stream));
    }
    p->src = static_cast<const uint8_t *>(c.stage_in.p);
    p->src_base = lo;
    return QD_OK;
}

static int ensure_twiddles(Chain &c, size_t W)
{
    if (c.twiddles_n == W) return QD_OK;
    std::vector<float> tw(4 * W, 0.0f); // w(W, .), then the same values in fk_stft's thread order
    fft_twiddles(W, tw.data());
    c.twiddles_packed = stft_thread_twiddles(W, tw.data(), tw.data() + 2 * W) != 0;
    QD_TRY(c.ensure(c.twiddles, 4 * W * sizeof(float)));
    QD_CUDA(cudaMemcpyAsync(c.twiddles.p, tw.data(), 4 * W * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    QD_CUDA(cudaStreamSynchronize(c.stream)); // tw is a stack-lifetime host vector
    c.twiddles_n = W;
    return QD_OK;
}

static int ensure_window(Chain &c, size_t W)
{
    if (c.window_n == W) return QD_OK;
    std::vector<float> w(W);
    blackman_harris(W, w.data());
    QD_TRY(c.ensure(c.window, W * sizeof(float)));
    QD_CUDA(cudaMemcpyAsync(c.window.p, w.data(), W * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    QD_CUDA(cudaStreamSynchronize(c.stream));
    c.window_n = W;
    return QD_OK;
}

static int copy_out(Chain &c, void *dst, const void *src, size_t bytes, int space)
{
    if (!bytes) return QD_OK;
    QD_CUDA(cudaMemcpyAsync(dst, src, bytes, space == QD_SPACE_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                            c.stream));
    return QD_OK;
}


void fill_fft_args(Chain &c, const SinkArgs &sink, size_t W, FftArgs *fa)
{
    memset(fa, 0, sizeof *fa);
    fa->in_pitch = W;
    fa->tw = static_cast<const float2 *>(c.twiddles.p);
    fa->twp = c.twiddles_packed && c.twiddles_n == W ? fa->tw + W : nullptr;
    fa->window = sink.windowed ? static_cast<const float *>(c.window.p) : nullptr;
    fa->W = static_cast<uint32_t>(W);
    fa->epi = sink.kind == SINK_SPARK ? EPI_SPARK : sink.kind == SINK_LEVELS ? EPI_LEVELS : EPI_TAKE;
    fa->mn = sink.min;
    fa->mx = sink.max;
    fa->distinction = (sink.max - sink.min) / 7.0f; // fft.rs:45, graph.len() == 7
    fa->panic_flag = static_cast<int *>(c.flag.p);
}

int launch_fft(Chain &c, FftArgs &fa, uint64_t units)
{
    if (units == 0) return QD_OK;
    const uint32_t W = fa.W;
    if (W & (W - 1)) { // not a power of two: direct DFT (take_fft only)
        if (fa.epi != EPI_TAKE || fa.raw) return set_error(QD_E_FFT_WIDTH, "internal: width %u needs the take_fft sink", W);
        fa.n_units = units;
        const size_t smem = static_cast<size_t>(W) * sizeof(float2);
        if (smem > 48 * 1024)
            QD_CUDA(cudaFuncSetAttribute(gk_dft, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        if (units > 0x7fffffffull) return set_error(QD_E_INVALID_ARG, "too many windows in one launch");
        gk_dft<<<static_cast<unsigned>(units), 256, smem, c.stream>>>(fa);
        QD_LAUNCHED();
        return QD_OK;
    }
    if (c.use_fast && fa.epi == EPI_SPARK && W <= 4096 && !fa.window) {
        bool handled = false;
        QD_TRY(launch_stft_fast(c, fa, units, &handled));
        if (handled) return QD_OK;
    }
    const uint32_t threads = 256;
    fa.team = std::min<uint32_t>(threads, std::max<uint32_t>(1, W / 4));
    const uint32_t wpc = threads / fa.team;
    fa.n_units = units;
    const size_t smem = static_cast<size_t>(wpc) * W * sizeof(float2);
    if (smem > 48 * 1024)
        QD_CUDA(cudaFuncSetAttribute(gk_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const uint64_t grid = (units + wpc - 1) / wpc;
    if (grid > 0x7fffffffull) return set_error(QD_E_INVALID_ARG, "too many windows in one launch");
    gk_fft<<<static_cast<unsigned>(grid), threads, smem, c.stream>>>(fa);
    QD_LAUNCHED();
    return QD_OK;
}

// element sizes of the sink's per-unit outputs
static size_t sink_idx_bytes(const SinkArgs &s, size_t W) { return s.kind == SINK_LEVELS ? 1 : (s.kind == SINK_SPARK ? W : 0); }
static bool sink_has_mag(const SinkArgs &s) { return s.kind == SINK_TAKE || (s.kind == SINK_SPARK && s.mag_out); }

// Sink processing of one fused-path segment: d_top holds [nu][unit_len] cf32 for units u0.. of the call.
struct FastSinkCtx {
    SinkArgs *sink;
    uint64_t unit_len;
    // windows cut straight from raw capture bytes (a chain without stages): see run_units_rawfft
    const uint8_t *raw = nullptr;
    int raw_fmt = 0;
    uint64_t raw_first = 0;
};

// STFT arguments of one fused-path segment: where its rows go (staging slot j for host sinks, the caller's device
// buffers otherwise) and the glyph thresholds.  Also the `prepare` callback of run_units_fast.
static int fast_sink_prepare(Chain &c, void *user, int j, uint64_t u0, uint64_t nu, FftArgs *fa)
{
    FastSinkCtx *ctx = static_cast<FastSinkCtx *>(user);
    SinkArgs &sink = *ctx->sink;
    const size_t W = sink.width;
    fill_fft_args(c, sink, W, fa);
    const size_t ib = sink_idx_bytes(sink, W);
    const bool mag = sink_has_mag(sink);
    if (sink.space == QD_SPACE_HOST) {
        if (ib) QD_TRY(c.ensure(c.pipe_idx[j], nu * ib));
        if (mag) QD_TRY(c.ensure(c.pipe_mag[j], nu * W * sizeof(float)));
        fa->idx = static_cast<uint8_t *>(c.pipe_idx[j].p);
        fa->mag = mag ? static_cast<float *>(c.pipe_mag[j].p) : nullptr;
    } else {
        fa->idx = ib ? sink.idx_out + u0 * ib : nullptr;
        fa->mag = mag ? sink.mag_out + u0 * W : nullptr;
    }
    fa->n_units = nu;
    stft_finalize_args(*fa);
    if (!c.glyph_lin) fa->use_lin = 0;
    return QD_OK;
}

static int fast_segment_sink(Chain &c, void *user, int j, uint64_t u0, uint64_t nu, const float2 *d_top, uint64_t pitch)
{
    FastSinkCtx *ctx = static_cast<FastSinkCtx *>(user);
    SinkArgs &sink = *ctx->sink;
    const size_t W = sink.width;
    const bool to_host = sink.space == QD_SPACE_HOST;
    if (sink.kind == SINK_SAMPLES) {
        if (!to_host) {
            // normally the kernel wrote straight into the caller's device buffer
            if (d_top != reinterpret_cast<const float2 *>(sink.samples_out) + u0 * ctx->unit_len)
                QD_CUDA(cudaMemcpyAsync(sink.samples_out + u0 * ctx->unit_len, d_top, nu * ctx->unit_len * sizeof(float2),
                                        cudaMemcpyDeviceToDevice, c.stream));
            return QD_OK;
        }
        QD_CUDA(cudaEventRecord(c.ev_sink[j], c.stream));
        QD_CUDA(cudaStreamWaitEvent(c.d2h_stream, c.ev_sink[j], 0));
        QD_CUDA(cudaMemcpyAsync(sink.samples_out + u0 * ctx->unit_len, d_top, nu * ctx->unit_len * sizeof(float2),
                                cudaMemcpyDeviceToHost, c.d2h_stream));
        QD_CUDA(cudaEventRecord(c.ev_d2h[j], c.d2h_stream));
        return QD_OK;
    }
    FftArgs fa;
    QD_TRY(fast_sink_prepare(c, user, j, u0, nu, &fa));
    const size_t ib = sink_idx_bytes(sink, W);
    const bool mag = sink_has_mag(sink);
    if (d_top || ctx->raw) { // (d_top == nullptr without raw windows: the filter kernel has already run the STFT)
        fa.in = d_top;
        fa.in_pitch = pitch; // unit_len for a [units][W] matrix, the window stride for a contiguous stream
        fa.tail = c.seg_tail; // truncated window tails patched over the stream (run_units_fast)
        fa.tail_len = c.seg_tail_len;
        fa.raw = ctx->raw;
        fa.raw_fmt = ctx->raw_fmt;
        fa.raw_first = ctx->raw_first;
        QD_TRY(launch_fft(c, fa, nu));
    }
    if (to_host) {
        QD_CUDA(cudaEventRecord(c.ev_sink[j], c.stream));
        QD_CUDA(cudaStreamWaitEvent(c.d2h_stream, c.ev_sink[j], 0));
        if (ib) QD_CUDA(cudaMemcpyAsync(sink.idx_out + u0 * ib, fa.idx, nu * ib, cudaMemcpyDeviceToHost, c.d2h_stream));
        if (mag)
            QD_CUDA(cudaMemcpyAsync(sink.mag_out + u0 * W, fa.mag, nu * W * sizeof(float), cudaMemcpyDeviceToHost, c.d2h_stream));
        QD_CUDA(cudaEventRecord(c.ev_d2h[j], c.d2h_stream));
    }
    return QD_OK;
}

// `from FILE | sparkfft`: no stage between the capture and the STFT.  Windows are decoded inside the FFT
// kernel's load, straight from the raw bytes (overlapping windows re-read them through L2); host and
// file sources are staged in double-buffered segments like the fused FIR path.
static int run_units_rawfft(Chain &c, uint64_t off0, uint64_t stride, uint64_t n_units, uint64_t W, SinkArgs &sink)
{
    const Source &s = c.src;
    const uint64_t pb = pair_bytes(s.format);
    const bool on_device = s.kind == QD_SRC_DEVICE_MEM;
    uint64_t seg_units = n_units;
    // a segment of nu windows stages (nu - 1) * stride + W samples: `stride` per window, also when stride > W
    if (!on_device) seg_units = std::max<uint64_t>(1, c.segment_bytes / std::max<uint64_t>(1, ((stride && n_units > 1) ? stride : W) * pb));
    if (sink.space == QD_SPACE_HOST)
        seg_units = std::min<uint64_t>(seg_units, std::max<uint64_t>(1, c.scratch_budget / 2 / (W * 5)));
    seg_units = std::min(seg_units, n_units);
    QD_TRY(c.ensure_pipeline());
    QD_CUDA(cudaEventRecord(c.ev_entry, c.stream));
    QD_CUDA(cudaStreamWaitEvent(c.h2d_stream, c.ev_entry, 0));
    QD_CUDA(cudaStreamWaitEvent(c.d2h_stream, c.ev_entry, 0));
    FastSinkCtx ctx{&sink, W};
    uint64_t seg = 0;
    for (uint64_t u0 = 0; u0 < n_units; u0 += seg_units, ++seg) {
        const int j = static_cast<int>(seg & 1);
        const uint64_t nu = std::min(seg_units, n_units - u0);
        const uint64_t lo = off0 + u0 * stride, hi = lo + (nu - 1) * stride + W;
        if (lo < s.base_sample || hi > s.base_sample + s.resident_samples)
            return set_error(QD_E_NOT_RESIDENT, "samples [%llu, %llu) requested but this source holds [%llu, %llu)",
                             (unsigned long long)lo, (unsigned long long)hi, (unsigned long long)s.base_sample,
                             (unsigned long long)(s.base_sample + s.resident_samples));
        if (on_device) {
            ctx.raw = s.data;
            ctx.raw_first = lo - s.base_sample;
        } else {
            const size_t bytes = static_cast<size_t>((hi - lo) * pb);
            QD_TRY(c.ensure(c.pipe_in[j], bytes + 64));
            if (seg >= 2) QD_CUDA(cudaStreamWaitEvent(c.h2d_stream, c.ev_compute[j], 0));
            if (s.kind == QD_SRC_HOST_MEM) {
                QD_CUDA(cudaMemcpyAsync(c.pipe_in[j].p, s.data + (lo - s.base_sample) * pb, bytes, cudaMemcpyHostToDevice,
                                        c.h2d_stream));
            } else {
                QD_TRY(c.ensure_pinned2(j, bytes));
                if (seg >= 2) QD_CUDA(cudaEventSynchronize(c.ev_h2d[j]));
                size_t done = 0;
                while (done < bytes) {
                    const ssize_t r = pread(s.fd, static_cast<uint8_t *>(c.h_pin2[j]) + done, bytes - done,
                                            static_cast<off_t>(lo * pb + done));
                    if (r <= 0) return set_error(QD_E_IO, "read %s: %s", s.path.c_str(), r < 0 ? strerror(errno) : "unexpected end of file");
                    done += static_cast<size_t>(r);
                }
                QD_CUDA(cudaMemcpyAsync(c.pipe_in[j].p, c.h_pin2[j], bytes, cudaMemcpyHostToDevice, c.h2d_stream));
            }
            QD_CUDA(cudaEventRecord(c.ev_h2d[j], c.h2d_stream));
            QD_CUDA(cudaStreamWaitEvent(c.stream, c.ev_h2d[j], 0));
            ctx.raw = static_cast<const uint8_t *>(c.pipe_in[j].p);
            ctx.raw_first = 0;
        }
        ctx.raw_fmt = s.format;
        if (seg >= 2) QD_CUDA(cudaStreamWaitEvent(c.stream, c.ev_d2h[j], 0));
        QD_TRY(c.prof_begin());
        QD_TRY(fast_segment_sink(c, &ctx, j, u0, nu, nullptr, stride));
        QD_TRY(c.prof_end("fk_stft (decode + STFT + magnitude + bucket)"));
        QD_CUDA(cudaEventRecord(c.ev_compute[j], c.stream));
    }
    return QD_OK;
}

// The unit-local path for units [u_begin, n_units) of the call.
static int run_units_generic(Chain &c, uint64_t off0, uint64_t stride, const uint64_t *offsets, uint64_t u_begin,
                             uint64_t n_units, uint64_t unit_len, SinkArgs &sink, uint64_t *produced_io)
{
    GPlan plan;
    QD_TRY(fill_plan(c, unit_len, &plan));
    const int S = plan.n_stages;
    const size_t W = sink.width;
    const bool fft_sink = sink.kind != SINK_SAMPLES;

    // leading shifts ride along with the source kernel
    int n_lead = 0;
    while (n_lead < S && plan.st[n_lead].kind == QD_STAGE_SHIFT) n_lead++;

    // scratch per unit: one buffer per level, shift levels alias their input level
    size_t per_unit = 0;
    for (int l = 0; l <= S; l++)
        if (l == 0 || plan.st[l - 1].kind == QD_STAGE_LOWPASS) per_unit += plan.n_level[l] * sizeof(float2);
    if (fft_sink) per_unit += W * (sizeof(uint8_t) + sizeof(float));
    uint64_t B = std::max<uint64_t>(1, c.scratch_budget / std::max<size_t>(per_unit, 1));
    B = std::min<uint64_t>(B, std::min<uint64_t>(n_units - u_begin, 32768));

    Chain::Buf *lvl[kMaxStages + 1];
    for (int l = 0; l <= S; l++) {
        if (l == 0 || plan.st[l - 1].kind == QD_STAGE_LOWPASS) {
            QD_TRY(c.ensure(c.level[l], B * plan.n_level[l] * sizeof(float2)));
            lvl[l] = &c.level[l];
        } else {
            lvl[l] = lvl[l - 1];
        }
    }
    QD_TRY(c.ensure(c.geo, B * (S + 1) * sizeof(uint32_t)));
    const size_t ib = sink_idx_bytes(sink, W);
    const bool mag = sink_has_mag(sink);
    const bool to_host = sink.space == QD_SPACE_HOST;
    if (fft_sink && to_host) {
        if (ib) QD_TRY(c.ensure(c.sink_a, B * ib));
        if (mag) QD_TRY(c.ensure(c.sink_b, B * W * sizeof(float)));
    }
    if (offsets) QD_TRY(c.ensure(c.offsets, B * sizeof(uint64_t)));

    uint64_t produced = *produced_io;
    std::vector<uint64_t> vcount;
    for (uint64_t u0 = u_begin; u0 < n_units; u0 += B) {
        const uint32_t b = static_cast<uint32_t>(std::min<uint64_t>(B, n_units - u0));
        const uint64_t boff = off0 + u0 * stride;

        // raw span of the batch
        uint64_t lo = UINT64_MAX, hi = 0;
        if (offsets) {
            for (uint32_t i = 0; i < b; i++) {
                uint64_t a, z;
                chain_source_span(c, offsets[u0 + i], unit_len, &a, &z);
                lo = std::min(lo, a);
                hi = std::max(hi, z);
            }
            QD_CUDA(cudaMemcpyAsync(c.offsets.p, offsets + u0, b * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
            QD_CUDA(cudaStreamSynchronize(c.stream));
        } else {
            uint64_t a, z;
            chain_source_span(c, boff, unit_len, &lo, &z);
            chain_source_span(c, boff + (b - 1) * stride, unit_len, &a, &hi);
            hi = std::max(hi, z);
        }
        QD_TRY(stage_source(c, lo, hi, &plan));
        const uint64_t *d_off = offsets ? static_cast<const uint64_t *>(c.offsets.p) : nullptr;
        uint32_t *d_valid = static_cast<uint32_t *>(c.geo.p);

        QD_TRY(c.prof_begin());
        gk_geometry<<<(b + 127) / 128, 128, 0, c.stream>>>(plan, boff, stride, d_off, b, d_valid);
        QD_LAUNCHED();
        {
            const uint32_t bpu = static_cast<uint32_t>((plan.n_level[0] + 255) / 256);
            gk_source<<<b * bpu, 256, 0, c.stream>>>(plan, n_lead, boff, stride, d_off, b, d_valid,
                                                     static_cast<float2 *>(lvl[0]->p), bpu);
            QD_LAUNCHED();
        }
        for (int s = n_lead; s < S; s++) {
            if (plan.st[s].kind == QD_STAGE_SHIFT) {
                const uint32_t bpu = static_cast<uint32_t>((plan.n_level[s + 1] + 255) / 256);
                gk_shift<<<b * bpu, 256, 0, c.stream>>>(plan, s, boff, stride, d_off, b, d_valid,
                                                        static_cast<float2 *>(lvl[s + 1]->p), bpu);
            } else {
                const uint32_t bpu = static_cast<uint32_t>((plan.n_level[s + 1] + 127) / 128);
                gk_lowpass<<<b * bpu, 128, 0, c.stream>>>(plan, s, b, d_valid, static_cast<const float2 *>(lvl[s]->p),
                                                          static_cast<float2 *>(lvl[s + 1]->p), bpu);
            }
            QD_LAUNCHED();
        }

        const float2 *top = static_cast<const float2 *>(lvl[S]->p);
        if (!fft_sink) {
            QD_TRY(c.prof_end("generic: gk_source+gk_shift+gk_lowpass"));
            // contiguous delivery of each unit's valid samples, runs of full units in one copy
            vcount.resize(b);
            for (uint32_t i = 0; i < b; i++) {
                const uint64_t o = offsets ? offsets[u0 + i] : boff + i * stride;
                QD_TRY(chain_valid(c, o, unit_len, &vcount[i]));
            }
            uint32_t i = 0;
            while (i < b) {
                uint32_t j = i;
                uint64_t n = 0;
                if (vcount[i] == unit_len) {
                    while (j < b && vcount[j] == unit_len) j++;
                    n = static_cast<uint64_t>(j - i) * unit_len;
                } else {
                    n = vcount[i];
                    j = i + 1;
                }
                QD_TRY(copy_out(c, sink.samples_out + produced, top + static_cast<size_t>(i) * unit_len,
                                n * sizeof(float2), sink.space));
                produced += n;
                i = j;
            }
        } else {
            FftArgs fa;
            fill_fft_args(c, sink, W, &fa);
            fa.in = top;
            if (to_host) {
                fa.idx = static_cast<uint8_t *>(c.sink_a.p);
                fa.mag = mag ? static_cast<float *>(c.sink_b.p) : nullptr;
            } else {
                fa.idx = ib ? sink.idx_out + u0 * ib : nullptr;
                fa.mag = mag ? sink.mag_out + u0 * W : nullptr;
            }
            QD_TRY(launch_fft(c, fa, b));
            QD_TRY(c.prof_end("generic: gk_source+gk_shift+gk_lowpass+gk_fft"));
            if (to_host) {
                if (ib) QD_TRY(copy_out(c, sink.idx_out + u0 * ib, fa.idx, static_cast<size_t>(b) * ib, sink.space));
                if (mag) QD_TRY(copy_out(c, sink.mag_out + u0 * W, fa.mag, static_cast<size_t>(b) * W * sizeof(float), sink.space));
            }
            produced += b;
        }
    }
    *produced_io = produced;
    return QD_OK;
}

int run_units(Chain &c, uint64_t off0, uint64_t stride, const uint64_t *offsets, uint64_t n_units, uint64_t unit_len,
              SinkArgs &sink, uint64_t *n_out)
{
    if (n_out) *n_out = 0;
    if (n_units == 0 || unit_len == 0) return QD_OK;
    QD_CUDA(cudaSetDevice(c.device));
    const size_t W = sink.width;
    const bool fft_sink = sink.kind != SINK_SAMPLES;
    if (fft_sink) {
        if (W != unit_len) return set_error(QD_E_INVALID_ARG, "internal: fft sink width != unit length");
        if (W * sizeof(float2) + 16 > 200 * 1024)
            return set_error(QD_E_INVALID_ARG, "fft width %zu exceeds the supported maximum of 16384", W);
        QD_TRY(ensure_twiddles(c, W));
        if (sink.windowed) QD_TRY(ensure_window(c, W));
        QD_TRY(c.ensure(c.flag, sizeof(int)));
        QD_CUDA(cudaMemsetAsync(c.flag.p, 0, sizeof(int), c.stream));
    }

    uint64_t produced = 0, done = 0;
    bool used_pipeline = false;
    if (!offsets && c.use_fast && fft_sink && c.stages.empty() && c.src.kind != QD_SRC_GEN) {
        QD_TRY(run_units_rawfft(c, off0, stride, n_units, unit_len, sink));
        done = produced = n_units;
        QD_CUDA(cudaEventRecord(c.ev_exit, c.d2h_stream));
        QD_CUDA(cudaStreamWaitEvent(c.stream, c.ev_exit, 0));
    } else if (!offsets && c.use_fast) {
        FastSinkCtx ctx{&sink, unit_len};
        float2 *direct = (sink.kind == SINK_SAMPLES && sink.space == QD_SPACE_DEVICE)
                             ? reinterpret_cast<float2 *>(sink.samples_out)
                             : nullptr;
        // the fast STFT kernel can take a window's truncated tail from a patch matrix (see run_units_fast)
        c.allow_tail = fft_sink && sink.kind == SINK_SPARK && !sink.windowed && unit_len <= 4096 && is_pow2(unit_len);
        // sparkfft can run inside the filter kernel (see run_units_fast)
        const bool can_fuse = sink.kind == SINK_SPARK && !sink.windowed;
        QD_TRY(run_units_fast(c, off0, stride, n_units, unit_len, direct, fast_segment_sink, &ctx, &done, can_fuse ? fast_sink_prepare : nullptr));
        c.allow_tail = false;
        used_pipeline = done > 0;
        produced = sink.kind == SINK_SAMPLES ? done * unit_len : done;
        if (used_pipeline) {
            // later work on the caller's stream (and the generic remainder) follows the drain copies
            QD_CUDA(cudaEventRecord(c.ev_exit, c.d2h_stream));
            QD_CUDA(cudaStreamWaitEvent(c.stream, c.ev_exit, 0));
        }
    }
    if (done < n_units) QD_TRY(run_units_generic(c, off0, stride, offsets, done, n_units, unit_len, sink, &produced));

    if (fft_sink && sink.kind == SINK_SPARK) {
        int flag = 0;
        QD_CUDA(cudaMemcpyAsync(&flag, c.flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        QD_CUDA(cudaStreamSynchronize(c.stream));
        sink.glyph_panic = flag != 0;
    } else if (sink.space == QD_SPACE_HOST) {
        QD_CUDA(cudaStreamSynchronize(c.stream));
    }
    if (n_out) *n_out = produced;
    return QD_OK;
}

} // namespace qd
