// qd_internal.h -- shared declarations of libquadrs_gpu (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <functional>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/quadrs_gpu.h"

struct qd_chain;

namespace qd {

// ---------------------------------------------------------------- errors
int set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
const char *last_error();
extern std::atomic<uint64_t> g_kernel_launches;

#define QD_CUDA(expr)                                                                                                  \
    do {                                                                                                               \
        cudaError_t e_ = (expr);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            return ::qd::set_error(QD_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__,        \
                                   __LINE__);                                                                          \
    } while (0)

#define QD_TRY(expr)                                                                                                   \
    do {                                                                                                               \
        int rc_ = (expr);                                                                                              \
        if (rc_ != QD_OK) return rc_;                                                                                  \
    } while (0)

// Count a launch of one of OUR kernels and surface launch errors.
#define QD_LAUNCHED()                                                                                                  \
    do {                                                                                                               \
        ::qd::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);                                               \
        QD_CUDA(cudaGetLastError());                                                                                   \
    } while (0)

// ---------------------------------------------------------------- host math (qd_host_math.cpp)
// All of these follow the reference's f32/f64 operation order exactly; see the .cpp for file:line.
uint64_t pair_bytes(int format);                                                   // lib.rs:217-229
double shift_ratio(int64_t frequency, uint64_t sample_rate);                       // shift.rs:28
void lowpass_taps(uint64_t frequency, uint64_t sample_rate, size_t size, float *out); // filter.rs:29-31,86-105
void blackman_harris(size_t n, float *out);                                        // ffts.rs:110-119
void fft_twiddles(size_t n, float *out_re_im);                                     // w(N,j), j<N (our FFT definition)
void sincos_table(double *out4x256);                                               // double-double cos/sin(2 pi i/256)
void sine_table_i16(int16_t *out4096);                                             // synthetic generator table
bool is_pow2(uint64_t v);
uint64_t f64_as_u64(double v);                                                     // Rust `as u64`

// ---------------------------------------------------------------- chain description
constexpr int kMaxStages = 8;
constexpr int kMaxTones = 16;

struct Stage {
    int kind = 0; // qd_stage_kind
    int64_t frequency = 0;
    uint64_t decimate = 1;
    uint64_t size = 0;    // taps
    uint64_t rate_in = 0; // sample rate of the stage's input
    double ratio = 0.0;   // shift
    std::vector<float> taps;
    float *d_taps = nullptr;
};

struct Source {
    int kind = 0;
    int format = 0;
    uint64_t sample_rate = 0;
    const uint8_t *data = nullptr; // host or device
    uint64_t n_bytes = 0;
    std::string path;
    int fd = -1;
    uint64_t base_sample = 0;
    uint64_t resident_samples = 0; // samples available at `data`
    uint64_t total_samples = 0;    // logical capture length
    double gen_seconds = 0;
    std::vector<int64_t> gen_cos;
};

// Per-device singletons (tables shared by all chains on a device).
struct DeviceCtx {
    int device = -1;
    double *d_sincos = nullptr; // 256 x {cos_hi, cos_lo, sin_hi, sin_lo}
    int16_t *d_sine_i16 = nullptr;
    int sm_count = 0;
    std::vector<int> local_cpus; // CPUs on the GPU's NUMA node (sysfs local_cpulist of its PCI function); may be empty
};
int device_ctx(int device, DeviceCtx **out);

// Plan handed to the generic ("unit-local") kernels by value.
struct GStage {
    int kind;
    uint32_t L;
    uint64_t D;
    double ratio;
    const float *taps;
};

struct GPlan {
    int n_stages;
    int src_kind;
    int fmt;
    int n_tones;
    GStage st[kMaxStages];
    uint64_t n_level[kMaxStages + 1]; // samples requested per unit at each level (0 = source)
    uint64_t mult[kMaxStages + 1];    // off_level = off_top * mult[level]
    const uint8_t *src;               // device pointer to sample `src_base`
    uint64_t src_base;
    uint64_t src_total;
    uint64_t src_rate;
    int64_t tones[kMaxTones];
    const double *sincos;
};

struct Chain {
    Source src;
    std::vector<Stage> stages;
    int device = 0;
    DeviceCtx *ctx = nullptr;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int precision = QD_PRECISION_EXACT;
    std::mutex mu;
    // qd_chain_create_sharded: the handle owns one complete chain per device and computes nothing itself
    std::vector<qd_chain *> shards;
    bool sharded() const { return !shards.empty(); }

    // device scratch (grown on demand, reused between calls)
    struct Buf {
        void *p = nullptr;
        size_t cap = 0;
    };
    Buf stage_in;             // uploaded raw bytes for HOST_MEM / FILE sources
    Buf level[kMaxStages + 1];
    Buf geo;
    Buf sink_a, sink_b, offsets;
    Buf twiddles, window;
    size_t twiddles_n = 0, window_n = 0;
    bool twiddles_packed = false; // the thread-ordered copy follows the natural table
    void *h_pinned = nullptr; // staging for FILE sources and small D2H
    size_t h_pinned_cap = 0;
    size_t scratch_budget = size_t(1) << 30;
    Buf flag; // glyph-panic flag of the spark sink

    // fused path (qd_fast.cu): segments are double buffered against H2D and D2H copies
    bool use_fast = true;
    int fuse_stft = 1; // sparkfft inside the filter kernel: 0 never, 1 for back-to-back windows (stride = width), 2 also for
                       // overlapping windows and two-stage chains (measured slower than the separate kernels: qd_fast.cu)
    // FAST arithmetic over cs8 captures: the filter runs on the tensor cores (fk_tcfir, qd_tcfir.cu) when the chain's
    // shape allows; 0 keeps the CUDA-core kernel
    int use_tc = 1;
    int stft_minb = 4; // experiment: resident CTAs per SM fk_stft<12> is compiled for (4: 64 registers, 3: 80, 2: 128)
    int glyph_lin = 1; // glyph indices through the linear form where it is proven (FftArgs::use_lin); 0 keeps the thresholds
    Buf tc_bimg;                  // B operand image of the current filter
    std::vector<uint8_t> tc_host; // its host copy (source of the upload)
    uint64_t tc_key[4] = {0, 0, 0, 0}; // (L, D, bits of the ratio sum, tap checksum) the image was built for
    float tc_s_hi = 0.0f, tc_s_lo = 0.0f;
    int fir_carry = 1;   // fk_fir: carry the overlap between a CTA's consecutive tiles (long filters); 0 = experiments
    int fir_cta_cap = 0; // experiments: resident fk_fir CTAs per SM (0 = as many as fit)
    size_t segment_bytes = size_t(32) << 20; // raw bytes staged per segment for host / file sources (measured best of 16..256 MiB)
    bool pipeline_ready = false;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_entry = nullptr, ev_exit = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_compute[2] = {nullptr, nullptr}, ev_sink[2] = {nullptr, nullptr},
                ev_d2h[2] = {nullptr, nullptr};
    Buf pipe_in[2], pipe_mid[2], pipe_out[2], pipe_idx[2], pipe_mag[2], pipe_tail[2];
    // overlapping windows behind a filter with truncated positions: the windows are cut from one stream and the
    // last `seg_tail_len` samples of every window come from a patch matrix [units][seg_tail_len] (run_units_fast
    // sets these for the sink of the current segment; allow_tail: the sink understands them)
    const float2 *seg_tail = nullptr;
    uint32_t seg_tail_len = 0;
    bool allow_tail = false;
    void *h_pin2[2] = {nullptr, nullptr};
    size_t h_pin2_cap[2] = {0, 0};
    int ensure_pipeline();
    int ensure_pinned2(int j, size_t bytes);

    // bench instrumentation (qd_chain_profile)
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    std::string prof_kernel;
    std::vector<std::string> prof_names;
    int prof_begin();
    int prof_end(const char *kernel);

    ~Chain();
    int ensure(Buf &b, size_t bytes);
    int ensure_pinned(size_t bytes);
};

} // namespace qd
struct qd_chain : qd::Chain {};
namespace qd {

// host-side cascade (exactly the reference's len()/read_at() count arithmetic)
int chain_len(const Chain &c, uint64_t *len);
uint64_t chain_rate(const Chain &c);
// samples a top-level read_at(off, n) returns; QD_E_* where the reference panics
int chain_valid(const Chain &c, uint64_t off, uint64_t n, uint64_t *valid);
// raw source sample range [lo, hi) a top-level read (off, n) touches (clamped to the capture)
void chain_source_span(const Chain &c, uint64_t off, uint64_t n, uint64_t *lo, uint64_t *hi);

// ---------------------------------------------------------------- executors (qd_generic.cu)
enum SinkKind { SINK_SAMPLES = 0, SINK_SPARK = 1, SINK_LEVELS = 2, SINK_TAKE = 3 };

struct SinkArgs {
    int kind = SINK_SAMPLES;
    size_t width = 0;
    float min = 0.08f, max = 1.0f;
    bool windowed = false;
    // outputs (host or device per `space`)
    int space = QD_SPACE_HOST;
    qd_cf32 *samples_out = nullptr; // SINK_SAMPLES: contiguous
    uint8_t *idx_out = nullptr;     // SINK_SPARK [units][width]; SINK_LEVELS [units]
    float *mag_out = nullptr;       // SINK_SPARK (nullable) / SINK_TAKE [units][width]
    bool glyph_panic = false;       // set when a bin hit the reference's graph[7] panic
};

// Runs `n_units` top-level reads of `unit_len` samples at offsets off0 + u*stride (or offsets[u])
// through the chain and the sink.  produced[] semantics: for SINK_SAMPLES, *n_out = total samples.
int run_units(Chain &c, uint64_t off0, uint64_t stride, const uint64_t *offsets, uint64_t n_units, uint64_t unit_len,
              SinkArgs &sink, uint64_t *n_out);

// Fused path (qd_fast.cu).  Handles the leading run of FULL units of the progression off0 + u*stride
// when the chain is From -> Shift* -> LowPass with a supported decimation; *units_done = how many.
// Output of segment units [u0, u0+nu) is [nu][unit_len] cf32 at d_direct + u0*unit_len when d_direct
// is given, else in a library staging buffer; on_segment (nullable) is called after each segment's
// kernel has been enqueued on c.stream, with j = staging slot.
// d_top/pitch: unit u of the segment starts at d_top + u*pitch (pitch = unit_len for a [units][n] matrix,
// = stride when the top stage was materialised as one contiguous stream).
struct FftArgs;
typedef int (*FastSegmentFn)(Chain &c, void *user, int j, uint64_t u0, uint64_t nu, const float2 *d_top, uint64_t pitch);
// prepare (nullable): the sink is sparkfft and can run inside the filter kernel when the units are back-to-back
// windows; it fills the STFT arguments (outputs of segment slot j, units u0..) and on_segment is then called with
// d_top == nullptr: the segment's rows are already in place, only the copies to the host remain.
typedef int (*FastPrepareFn)(Chain &c, void *user, int j, uint64_t u0, uint64_t nu, FftArgs *fa);
int run_units_fast(Chain &c, uint64_t off0, uint64_t stride, uint64_t n_units, uint64_t unit_len, float2 *d_direct, FastSegmentFn on_segment, void *user, uint64_t *units_done, FastPrepareFn prepare = nullptr);

// ---------------------------------------------------------------- STFT kernels (qd_generic.cu, qd_stft.cu)
enum { EPI_SPARK = 0, EPI_LEVELS = 1, EPI_TAKE = 2 };

struct FftArgs {
    const float2 *in;     // cf32 windows: window u starts at in + u * in_pitch
    const float2 *tail;   // nullable: the last tail_len samples of window u are tail[u * tail_len ..] instead
    uint32_t tail_len;
    uint64_t in_pitch;
    const uint8_t *raw;   // or raw capture bytes (decoded on load): window u starts at sample raw_first + u * in_pitch
    int raw_fmt;
    uint64_t raw_first;
    uint64_t n_units;
    const float2 *tw;     // w(W, j), j < W
    const float2 *twp;    // nullable: the same values in the order fk_stft's threads use them (stft_thread_twiddles)
    const float *window;  // nullable (take_fft BlackmanHarris)
    uint32_t W;
    uint32_t team;        // threads cooperating on one window
    int epi;
    float mn, mx, distinction;
    uint8_t *idx;         // SPARK [units][W]; LEVELS [units]
    float *mag;           // SPARK nullable / TAKE [units][W]
    int *panic_flag;
    // glyph boundaries as thresholds on re^2 + im^2 in f64 (spark_thresholds): lets the index be found
    // without the square root when magnitudes are not requested
    int use_thr;
    double thr[9];
    uint32_t thr_hi[9];   // their high words: s >= thr decides on the high word alone unless the two are equal
    // The same boundaries as one linear function of the square root, g = lin_a * sqrt(s) + lin_b (glyph = floor g
    // clamped to 0..8): evaluated in f32 and trusted only where g is further than lin_eps from an integer, every
    // other bin goes through the thresholds (stft_finalize_args proves the margin on the host, else use_lin = 0).
    int use_lin;
    float lin_a, lin_bh, lin_bl; // slope; lin_b - 0.5 + lin_eps and lin_b - 0.5 - lin_eps
    float lin_lo, lin_hi;        // clamp of the square root: g stays inside [0.5, 8.5]
    float2 one;           // (1, 1), opaque to ptxas (see qd_stft.cu pmul_tw)
};
// thr[c], c = 0..6: smallest s = fl64(re^2 + im^2) whose glyph index is >= c + 1; thr[7]: start of the
// graph[7] panic zone; thr[8]: smallest s with norm >= max.  false when min/max make that ill-defined.
bool spark_thresholds(float mn, float mx, double thr[9]);
int launch_stft_fast(Chain &c, const FftArgs &fa, uint64_t units, bool *handled);
void stft_finalize_args(FftArgs &fa);
// fk_stft's 16-point passes: pass p (q = 2^first_bits * 16^p points combined so far) reads its 15 twiddles per
// thread as out[15 * (sum of earlier q) + j * q + k]; returns the number of float2 written (<= W; 0: no such pass)
size_t stft_thread_twiddles(size_t W, const float *tw_re_im, float *out_re_im);

// ---------------------------------------------------------------- single-process multi-GPU (qd_multi.cu)
// Runs fn(i, shard i) for every shard of a sharded chain, each on its own host thread bound to the CPUs local
// to the shard's GPU; rc[i] / msg[i] receive each call's status and thread-local error text.
void run_on_shards(Chain &c, const std::function<int(size_t, qd_chain *)> &fn, std::vector<int> &rc, std::vector<std::string> &msg);
// status of the first failing shard in shard order (its message becomes the caller's qd_last_error), or QD_OK
int first_shard_error(const std::vector<int> &rc, const std::vector<std::string> &msg, size_t *which);
// [begin, end) of part i when n units are cut into parts contiguous ranges
inline void shard_range(uint64_t n, size_t parts, size_t i, uint64_t *begin, uint64_t *end)
{
    *begin = static_cast<uint64_t>((static_cast<unsigned __int128>(n) * i) / parts);
    *end = static_cast<uint64_t>((static_cast<unsigned __int128>(n) * (i + 1)) / parts);
}
void bind_thread_to_device_cpus(const DeviceCtx *ctx);
std::vector<int> device_local_cpus(int device);

int synth_fill(const qd_synth *p, int format, uint64_t first, uint64_t n, void *d_out, int device, cudaStream_t st);

} // namespace qd
