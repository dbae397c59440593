// qd_capi.cu -- the extern "C" entry points declared in include/quadrs_gpu.h.
//
// Host logic here mirrors the reference's control flow around the hot path: Operation::exec's
// construction checks (lib.rs:89-121), the sink loops of fft.rs / ffts.rs / lib.rs do_write, and
// their panic / Err conditions mapped to status codes.  All sample arithmetic runs on the device;
// there is no CPU fallback anywhere in this library.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "qd_internal.h"

namespace qd {

static thread_local char tl_error[512] = "";
std::atomic<uint64_t> g_kernel_launches{0};

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof tl_error, fmt, ap);
    va_end(ap);
    return code;
}

const char *last_error() { return tl_error; }

static std::mutex g_ctx_mu;
static std::map<int, std::unique_ptr<DeviceCtx>> g_ctx;

int device_ctx(int device, DeviceCtx **out)
{
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) {
        *out = it->second.get();
        return QD_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_error(QD_E_CUDA, "no usable CUDA device (%s); libquadrs_gpu has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return set_error(QD_E_INVALID_ARG, "device %d out of range (%d devices)", device, count);
    QD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(QD_E_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                         prop.minor);
    auto ctx = std::make_unique<DeviceCtx>();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->local_cpus = device_local_cpus(device);
    double tab[4 * 256];
    sincos_table(tab);
    QD_CUDA(cudaMalloc(&ctx->d_sincos, sizeof tab));
    QD_CUDA(cudaMemcpy(ctx->d_sincos, tab, sizeof tab, cudaMemcpyHostToDevice));
    int16_t sine[4096];
    sine_table_i16(sine);
    QD_CUDA(cudaMalloc(&ctx->d_sine_i16, sizeof sine));
    QD_CUDA(cudaMemcpy(ctx->d_sine_i16, sine, sizeof sine, cudaMemcpyHostToDevice));
    *out = ctx.get();
    g_ctx[device] = std::move(ctx);
    return QD_OK;
}

// Fills the host-side description (no device work): shared by qd_chain_create and qd_shard_plan.
static int describe(const qd_source *src, const qd_stage *stages, size_t n_stages, Chain &c)
{
    if (!src) return set_error(QD_E_INVALID_ARG, "source is null");
    if (n_stages > static_cast<size_t>(kMaxStages)) return set_error(QD_E_INVALID_ARG, "at most %d stages", kMaxStages);
    if (n_stages && !stages) return set_error(QD_E_INVALID_ARG, "stages is null");
    Source &s = c.src;
    s.kind = src->kind;
    s.format = src->format;
    s.sample_rate = src->sample_rate;
    s.base_sample = src->base_sample;
    if (s.kind == QD_SRC_GEN) {
        // Gen::new, gen.rs:17-21
        if (src->gen_n_cos == 0 || !src->gen_cos) return set_error(QD_E_GEN_ARGS, "cos cannot be empty");
        if (src->gen_n_cos > static_cast<uint64_t>(kMaxTones)) return set_error(QD_E_INVALID_ARG, "at most %d tones", kMaxTones);
        if (src->sample_rate == 0) return set_error(QD_E_GEN_ARGS, "sample rate may not be zero");
        if (!(src->gen_seconds > 0.0)) return set_error(QD_E_GEN_ARGS, "seconds may not be <= 0");
        s.gen_seconds = src->gen_seconds;
        s.gen_cos.assign(src->gen_cos, src->gen_cos + src->gen_n_cos);
        s.total_samples = f64_as_u64(s.gen_seconds * static_cast<double>(s.sample_rate));
    } else {
        const uint64_t pb = pair_bytes(s.format);
        if (!pb) return set_error(QD_E_INVALID_ARG, "unknown sample format %d", s.format);
        if (s.kind == QD_SRC_FILE) {
            if (!src->path) return set_error(QD_E_INVALID_ARG, "file source without a path");
            s.path = src->path;
            s.fd = open(src->path, O_RDONLY);
            if (s.fd < 0) return set_error(QD_E_IO, "%s: %s", src->path, strerror(errno));
            const off_t end = lseek(s.fd, 0, SEEK_END); // SampleFile::new, samples.rs:52
            if (end < 0) return set_error(QD_E_IO, "seeking to end of %s: %s", src->path, strerror(errno));
            s.n_bytes = src->n_bytes ? std::min<uint64_t>(src->n_bytes, static_cast<uint64_t>(end)) : static_cast<uint64_t>(end);
        } else if (s.kind == QD_SRC_HOST_MEM || s.kind == QD_SRC_DEVICE_MEM) {
            if (!src->data && src->n_bytes) return set_error(QD_E_INVALID_ARG, "memory source without data");
            s.data = static_cast<const uint8_t *>(src->data);
            s.n_bytes = src->n_bytes;
            if (s.kind == QD_SRC_DEVICE_MEM && reinterpret_cast<uintptr_t>(src->data) % pb != 0)
                return set_error(QD_E_INVALID_ARG, "device capture pointer must be aligned to one sample (%llu bytes)",
                                 (unsigned long long)pb);
        } else {
            return set_error(QD_E_INVALID_ARG, "unknown source kind %d", s.kind);
        }
        s.resident_samples = s.n_bytes / pb; // samples.rs:64-66: trailing partial pair is not a sample
        s.total_samples = src->total_samples ? src->total_samples : s.base_sample + s.resident_samples;
        if (s.base_sample + s.resident_samples > s.total_samples)
            return set_error(QD_E_INVALID_ARG, "shard [%llu, +%llu) exceeds total_samples %llu",
                             (unsigned long long)s.base_sample, (unsigned long long)s.resident_samples,
                             (unsigned long long)s.total_samples);
    }
    uint64_t rate = s.sample_rate;
    for (size_t i = 0; i < n_stages; i++) {
        Stage st;
        st.kind = stages[i].kind;
        st.frequency = stages[i].frequency;
        st.rate_in = rate;
        if (st.kind == QD_STAGE_SHIFT) {
            // Shift::new, shift.rs:19-31 (asserts in the reference's order)
            const int64_t a = st.frequency < 0 ? -st.frequency : st.frequency;
            if (!(a < static_cast<int64_t>(rate / 2)))
                return set_error(QD_E_SHIFT_NYQUIST, "frequency must be under half the sample rate");
            if (!(rate > 0)) return set_error(QD_E_ZERO_RATE, "assertion failed: sample_rate > 0");
            st.ratio = shift_ratio(st.frequency, rate);
        } else if (st.kind == QD_STAGE_LOWPASS) {
            st.decimate = stages[i].decimate;
            st.size = stages[i].size;
            if (st.decimate == 0) return set_error(QD_E_INVALID_ARG, "lowpass: decimate 0 (the reference divides by zero)");
            if (st.size == 0 || st.size >= (uint64_t(1) << 24))
                return set_error(QD_E_INVALID_ARG, "lowpass: filter size %llu unsupported", (unsigned long long)st.size);
            if (st.frequency < 0) return set_error(QD_E_INVALID_ARG, "lowpass: frequency is unsigned (lib.rs:37)");
            st.taps.resize(st.size);
            lowpass_taps(static_cast<uint64_t>(st.frequency), rate, st.size, st.taps.data());
            rate = rate / st.decimate; // filter.rs:50-52
        } else {
            return set_error(QD_E_INVALID_ARG, "unknown stage kind %d", st.kind);
        }
        c.stages.push_back(std::move(st));
    }
    return QD_OK;
}

static int first_bad_unit(const Chain &c, uint64_t off0, uint64_t stride, uint64_t n_units, uint64_t unit_len,
                          uint64_t *n_good, int *bad_rc)
{
    // valid(off) is non-increasing in off and panics only appear at larger offsets: bisect.
    auto full = [&](uint64_t u, int *rc) {
        uint64_t v = 0;
        const int r = chain_valid(c, off0 + u * stride, unit_len, &v);
        *rc = r != QD_OK ? r : (v == unit_len ? QD_OK : QD_E_SHORT_READ);
        return *rc == QD_OK;
    };
    int rc;
    *bad_rc = QD_OK;
    if (n_units == 0 || full(n_units - 1, &rc)) {
        *n_good = n_units;
        return QD_OK;
    }
    uint64_t lo = 0, hi = n_units - 1; // hi is bad
    while (lo < hi) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (full(mid, &rc)) lo = mid + 1;
        else hi = mid;
    }
    full(lo, &rc);
    if (rc == QD_E_SHORT_READ) {
        uint64_t v = 0;
        chain_valid(c, off0 + lo * stride, unit_len, &v);
        set_error(QD_E_SHORT_READ, "TODO: read-exact messed up: %llu (wanted) != %llu (read) at %llu",
                  (unsigned long long)unit_len, (unsigned long long)v, (unsigned long long)(off0 + lo * stride));
    }
    *n_good = lo;
    *bad_rc = rc;
    return QD_OK;
}

} // namespace qd

using namespace qd;

// applies one call to every shard of a sharded handle (settings, synchronisation)
template <class F>
static int each_shard(qd_chain *c, F fn)
{
    for (qd_chain *s : c->shards) QD_TRY(fn(s));
    return QD_OK;
}

extern "C" {

const char *qd_last_error(void) { return last_error(); }
int qd_abi_version(void) { return QD_ABI_VERSION; }
uint64_t qd_kernel_launches(void) { return g_kernel_launches.load(); }

const char *qd_status_name(int s)
{
    static const char *names[] = {"QD_OK", "QD_E_INVALID_ARG", "QD_E_SHIFT_NYQUIST", "QD_E_ZERO_RATE", "QD_E_OFFSET_EOF",
                                  "QD_E_SHORT_INPUT", "QD_E_SHORT_READ", "QD_E_FFT_WIDTH", "QD_E_GLYPH_RANGE", "QD_E_LEVELS",
                                  "QD_E_SLICE", "QD_E_VISIBLE", "QD_E_GEN_ARGS", "QD_E_WRITE_SHORT", "QD_E_IO", "QD_E_CUDA",
                                  "QD_E_NOT_RESIDENT", "QD_E_UNIMPLEMENTED", "QD_E_EXISTS", "QD_E_NOMEM", "QD_E_ZERO_STRIDE"};
    return s >= 0 && s <= 20 ? names[s] : "QD_E_UNKNOWN";
}

int qd_device_count(int *count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        if (count) *count = 0;
        return set_error(QD_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    if (count) *count = n;
    return QD_OK;
}

int qd_chain_create(const qd_source *src, const qd_stage *stages, size_t n_stages, int device, qd_chain **out)
{
    if (!out) return set_error(QD_E_INVALID_ARG, "out is null");
    *out = nullptr;
    std::unique_ptr<qd_chain> c(new qd_chain());
    c->device = device;
    QD_TRY(describe(src, stages, n_stages, *c));
    QD_TRY(device_ctx(device, &c->ctx));
    QD_CUDA(cudaSetDevice(device));
    QD_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    for (Stage &s : c->stages) {
        if (s.kind != QD_STAGE_LOWPASS) continue;
        QD_CUDA(cudaMalloc(&s.d_taps, s.size * sizeof(float)));
        QD_CUDA(cudaMemcpy(s.d_taps, s.taps.data(), s.size * sizeof(float), cudaMemcpyHostToDevice));
    }
    *out = c.release();
    return QD_OK;
}

int qd_chain_create_sharded(const qd_source *src, const qd_stage *stages, size_t n_stages, const int *devices,
                            size_t n_dev, qd_chain **out)
{
    if (!out) return set_error(QD_E_INVALID_ARG, "out is null");
    *out = nullptr;
    if (!devices || n_dev == 0) return set_error(QD_E_INVALID_ARG, "no devices given");
    if (n_dev > 64) return set_error(QD_E_INVALID_ARG, "at most 64 devices");
    if (n_dev == 1) return qd_chain_create(src, stages, n_stages, devices[0], out);
    if (!src) return set_error(QD_E_INVALID_ARG, "source is null");
    if (src->kind == QD_SRC_DEVICE_MEM)
        return set_error(QD_E_INVALID_ARG, "a DEVICE_MEM capture lives on one device: shard a HOST_MEM, FILE or GEN source, "
                                           "or create one chain per device over its own resident range");
    std::unique_ptr<qd_chain> c(new qd_chain());
    c->device = devices[0];
    for (size_t i = 0; i < n_dev; i++) {
        qd_chain *child = nullptr;
        const int rc = qd_chain_create(src, stages, n_stages, devices[i], &child);
        if (rc != QD_OK) return rc; // ~Chain of the handle destroys the children made so far
        c->shards.push_back(child);
    }
    *out = c.release();
    return QD_OK;
}

int qd_chain_n_devices(const qd_chain *c, size_t *n_dev)
{
    if (!c || !n_dev) return set_error(QD_E_INVALID_ARG, "null argument");
    *n_dev = c->sharded() ? c->shards.size() : 1;
    return QD_OK;
}

void qd_chain_destroy(qd_chain *c) { delete c; }


int qd_chain_set_stream(qd_chain *c, void *cuda_stream)
{
    if (!c) return set_error(QD_E_INVALID_ARG, "chain is null");
    if (c->sharded()) return set_error(QD_E_INVALID_ARG, "a sharded chain runs on its own per-device streams");
    std::lock_guard<std::mutex> lk(c->mu);
    QD_CUDA(cudaSetDevice(c->device));
    QD_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream) QD_CUDA(cudaStreamDestroy(c->stream));
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    c->own_stream = false;
    return QD_OK;
}

int qd_chain_set_precision(qd_chain *c, int precision)
{
    if (!c) return set_error(QD_E_INVALID_ARG, "chain is null");
    if (precision != QD_PRECISION_EXACT && precision != QD_PRECISION_FAST)
        return set_error(QD_E_INVALID_ARG, "unknown precision %d", precision);
    if (c->sharded()) return each_shard(c, [&](qd_chain *s) { return qd_chain_set_precision(s, precision); });
    if (precision == QD_PRECISION_FAST && c->src.kind != QD_SRC_GEN &&
        (c->src.format == QD_FMT_CU8 || c->src.format == QD_FMT_CS16))
        return set_error(QD_E_INVALID_ARG,
                         "FAST precision is refused for cu8/cs16: the reference's own f32 rounding noise on these "
                         "offset formats exceeds 1e-5, so only its exact operation order reproduces it");
    std::lock_guard<std::mutex> lk(c->mu);
    c->precision = precision;
    return QD_OK;
}

int qd_chain_set_option(qd_chain *c, const char *key, int64_t value)
{
    if (!c || !key) return set_error(QD_E_INVALID_ARG, "null argument");
    if (c->sharded()) return each_shard(c, [&](qd_chain *s) { return qd_chain_set_option(s, key, value); });
    std::lock_guard<std::mutex> lk(c->mu);
    if (!strcmp(key, "use_fast")) c->use_fast = value != 0;
    else if (!strcmp(key, "fuse_stft") && value >= 0 && value <= 2) c->fuse_stft = static_cast<int>(value);
    else if (!strcmp(key, "use_tc") && value >= 0 && value <= 1) c->use_tc = static_cast<int>(value);
    else if (!strcmp(key, "glyph_lin") && value >= 0 && value <= 1) c->glyph_lin = static_cast<int>(value);
    else if (!strcmp(key, "stft_minb") && value >= 2 && value <= 4) c->stft_minb = static_cast<int>(value);
    else if (!strcmp(key, "fir_cta_cap") && value >= 0) c->fir_cta_cap = static_cast<int>(value);
    else if (!strcmp(key, "fir_carry") && value >= 0 && value <= 1) c->fir_carry = static_cast<int>(value);
    else if (!strcmp(key, "segment_bytes") && value > 0) c->segment_bytes = static_cast<size_t>(value);
    else if (!strcmp(key, "scratch_budget") && value > 0) c->scratch_budget = static_cast<size_t>(value);
    else return set_error(QD_E_INVALID_ARG, "unknown option %s=%lld", key, (long long)value);
    return QD_OK;
}

int qd_chain_synchronize(qd_chain *c)
{
    if (!c) return set_error(QD_E_INVALID_ARG, "chain is null");
    if (c->sharded()) return each_shard(c, [](qd_chain *s) { return qd_chain_synchronize(s); });
    QD_CUDA(cudaSetDevice(c->device));
    QD_CUDA(cudaStreamSynchronize(c->stream));
    return QD_OK;
}

int qd_chain_profile(qd_chain *c, int enable)
{
    if (!c) return set_error(QD_E_INVALID_ARG, "chain is null");
    if (c->sharded()) return each_shard(c, [&](qd_chain *s) { return qd_chain_profile(s, enable); });
    std::lock_guard<std::mutex> lk(c->mu);
    QD_CUDA(cudaSetDevice(c->device));
    QD_CUDA(cudaStreamSynchronize(c->stream));
    c->profile = enable != 0;
    c->prof_used = 0;
    return QD_OK;
}

int qd_chain_profile_read(qd_chain *c, uint64_t *regions, double *total_ms, char *kernel_name, size_t cap)
{
    if (!c) return set_error(QD_E_INVALID_ARG, "chain is null");
    if (c->sharded()) { // the devices run side by side: report the slowest one
        uint64_t n_max = 0;
        double ms_max = -1.0;
        for (qd_chain *s : c->shards) {
            uint64_t n = 0;
            double ms = 0.0;
            char name[256] = "";
            QD_TRY(qd_chain_profile_read(s, &n, &ms, name, sizeof name));
            if (ms > ms_max) {
                ms_max = ms;
                n_max = n;
                if (kernel_name && cap) snprintf(kernel_name, cap, "%s", name);
            }
        }
        if (regions) *regions = n_max;
        if (total_ms) *total_ms = ms_max < 0 ? 0.0 : ms_max;
        return QD_OK;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    QD_CUDA(cudaSetDevice(c->device));
    QD_CUDA(cudaStreamSynchronize(c->stream));
    // report the kernel (region name) with the largest summed device time, and that sum
    std::map<std::string, std::pair<double, uint64_t>> by_name;
    for (size_t i = 0; i < c->prof_used; i++) {
        float t = 0.0f;
        QD_CUDA(cudaEventElapsedTime(&t, c->prof_events[i].first, c->prof_events[i].second));
        auto &e = by_name[c->prof_names[i]];
        e.first += t;
        e.second += 1;
    }
    double ms = 0.0;
    uint64_t n = 0;
    for (auto &kv : by_name)
        if (kv.second.first > ms) {
            ms = kv.second.first;
            n = kv.second.second;
            c->prof_kernel = kv.first;
        }
    if (regions) *regions = n;
    if (total_ms) *total_ms = ms;
    if (kernel_name && cap) snprintf(kernel_name, cap, "%s", c->prof_kernel.c_str());
    c->prof_used = 0;
    return QD_OK;
}

int qd_chain_len(const qd_chain *c, uint64_t *len)
{
    if (!c || !len) return set_error(QD_E_INVALID_ARG, "null argument");
    if (c->sharded()) return chain_len(*c->shards[0], len);
    return chain_len(*c, len);
}

int qd_chain_sample_rate(const qd_chain *c, uint64_t *rate)
{
    if (!c || !rate) return set_error(QD_E_INVALID_ARG, "null argument");
    if (c->sharded()) c = c->shards[0];
    *rate = chain_rate(*c);
    return QD_OK;
}

int qd_chain_taps(const qd_chain *c, size_t stage, float *out, size_t cap, size_t *n)
{
    if (c && c->sharded()) c = c->shards[0];
    if (!c || stage >= c->stages.size() || c->stages[stage].kind != QD_STAGE_LOWPASS)
        return set_error(QD_E_INVALID_ARG, "stage %zu is not a lowpass", stage);
    const auto &t = c->stages[stage].taps;
    if (n) *n = t.size();
    if (out) memcpy(out, t.data(), std::min(cap, t.size()) * sizeof(float));
    return QD_OK;
}

int qd_chain_read_at(qd_chain *c, uint64_t off, qd_cf32 *buf, size_t n, int space, size_t *got)
{
    if (!c || (!buf && n) || !got) return set_error(QD_E_INVALID_ARG, "null argument");
    *got = 0;
    if (n == 0) return QD_OK;
    if (c->sharded()) return qd_chain_read_at(c->shards[0], off, buf, n, space, got); // one unit: one device
    std::lock_guard<std::mutex> lk(c->mu);
    uint64_t v = 0;
    QD_TRY(chain_valid(*c, off, n, &v));
    SinkArgs sink;
    sink.kind = SINK_SAMPLES;
    sink.space = space;
    sink.samples_out = buf;
    uint64_t produced = 0;
    QD_TRY(run_units(*c, off, 0, nullptr, 1, n, sink, &produced));
    *got = static_cast<size_t>(produced);
    return QD_OK;
}

int qd_chain_read_exact_at(qd_chain *c, uint64_t off, qd_cf32 *buf, size_t n, int space)
{
    size_t got = 0;
    QD_TRY(qd_chain_read_at(c, off, buf, n, space, &got));
    if (got != n) // samples.rs:19-25
        return set_error(QD_E_SHORT_READ, "TODO: read-exact messed up: %zu (wanted) != %zu (read) at %llu", n, got,
                         (unsigned long long)off);
    return QD_OK;
}

int qd_sparkfft_rows(const qd_chain *c, size_t width, uint64_t stride, uint64_t *rows)
{
    if (!c || !rows) return set_error(QD_E_INVALID_ARG, "null argument");
    if (stride == 0) return set_error(QD_E_ZERO_STRIDE, "stride 0 never terminates (fft.rs:65)");
    if (c->sharded()) c = c->shards[0];
    uint64_t len = 0;
    QD_TRY(chain_len(*c, &len));
    // while i < len - width { ...; i += stride }  (fft.rs:27-28,65)
    *rows = len > width ? (len - width + stride - 1) / stride : 0;
    return QD_OK;
}

int qd_sparkfft(qd_chain *c, size_t width, uint64_t stride, int has_range, float min, float max, uint64_t first_row,
                uint64_t n_rows, uint8_t *idx_out, float *mag_out, int space, uint64_t *rows_out)
{
    if (!c || !rows_out) return set_error(QD_E_INVALID_ARG, "null argument");
    *rows_out = 0;
    if (!is_pow2(width)) // Radix4::new, fft.rs:25
        return set_error(QD_E_FFT_WIDTH, "Radix4 algorithm requires a power-of-two input size. Got %zu", width);
    if (stride == 0) return set_error(QD_E_ZERO_STRIDE, "stride 0 never terminates (fft.rs:65)");
    if (c->sharded()) {
        // rows [first_row, first_row + n) in contiguous parts, one per device, each into its place of idx_out
        if (space != QD_SPACE_HOST) return set_error(QD_E_INVALID_ARG, "a sharded chain delivers into host buffers");
        uint64_t len = 0;
        QD_TRY(chain_len(*c->shards[0], &len));
        const uint64_t total = len >= width ? (len > width ? (len - width + stride - 1) / stride : 0) : 1;
        if (first_row >= total || n_rows == 0) return QD_OK;
        const uint64_t n = std::min(n_rows, total - first_row);
        if (!idx_out) return set_error(QD_E_INVALID_ARG, "idx_out is null");
        const size_t parts = c->shards.size();
        std::vector<uint64_t> got(parts, 0);
        std::vector<int> rc;
        std::vector<std::string> msg;
        run_on_shards(*c, [&](size_t i, qd_chain *s) {
            uint64_t a, b;
            shard_range(n, parts, i, &a, &b);
            if (b == a) return static_cast<int>(QD_OK);
            return qd_sparkfft(s, width, stride, has_range, min, max, first_row + a, b - a, idx_out + a * width,
                               mag_out ? mag_out + a * width : nullptr, space, &got[i]);
        }, rc, msg);
        // the reference stops at its first failing row: rows of later parts are not delivered
        size_t bad = parts;
        const int status = first_shard_error(rc, msg, &bad);
        for (size_t i = 0; i < parts && i <= bad; i++) *rows_out += got[i];
        return status;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    uint64_t len = 0;
    QD_TRY(chain_len(*c, &len));
    uint64_t total;
    if (len >= width) {
        total = len > width ? (len - width + stride - 1) / stride : 0;
    } else {
        total = 1; // len - width wraps (release build): the loop runs and its first read_exact_at fails
    }
    if (first_row >= total || n_rows == 0) return QD_OK;
    uint64_t n = std::min(n_rows, total - first_row);
    if (!idx_out) return set_error(QD_E_INVALID_ARG, "idx_out is null");
    uint64_t good = 0;
    int bad = QD_OK;
    QD_TRY(first_bad_unit(*c, first_row * stride, stride, n, width, &good, &bad));
    std::string bad_msg = bad != QD_OK ? last_error() : "";
    SinkArgs sink;
    sink.kind = SINK_SPARK;
    sink.width = width;
    sink.min = has_range ? min : 0.08f; // fft.rs:22-23
    sink.max = has_range ? max : 1.0f;
    sink.space = space;
    sink.idx_out = idx_out;
    sink.mag_out = mag_out;
    uint64_t produced = 0;
    QD_TRY(run_units(*c, first_row * stride, stride, nullptr, good, width, sink, &produced));
    *rows_out = produced;
    if (sink.glyph_panic)
        return set_error(QD_E_GLYPH_RANGE, "index out of bounds: the len is 7 but the index is 7 (fft.rs:59); bins "
                                           "that hit it are marked 9 in idx_out");
    if (bad != QD_OK) return set_error(bad, "%s", bad_msg.c_str());
    return QD_OK;
}

static const char *const kGlyphs[10] = {" ", "▁", "▂", "▃", "▄", "▅", "▆", "▇", "█", "?"};

size_t qd_format_row(const uint8_t *idx, size_t width, char *out, size_t cap)
{
    // fft.rs:34-36,63: "│" + glyphs + "│"
    size_t o = 0;
    auto put = [&](const char *s) {
        const size_t l = strlen(s);
        if (o + l <= cap) memcpy(out + o, s, l);
        o += l;
    };
    put("│");
    for (size_t b = 0; b < width; b++) put(kGlyphs[idx[b] <= 9 ? idx[b] : 9]);
    put("│");
    return o;
}

int qd_freq_levels(qd_chain *c, size_t width, uint64_t stride, size_t levels, uint64_t first, uint64_t n,
                   uint8_t *vals, int space, uint64_t *total_out)
{
    if (!c || !total_out) return set_error(QD_E_INVALID_ARG, "null argument");
    *total_out = 0;
    if (levels != 2) return set_error(QD_E_LEVELS, "only supporting two levels for now"); // fft.rs:83
    if (!is_pow2(width))
        return set_error(QD_E_FFT_WIDTH, "Radix4 algorithm requires a power-of-two input size. Got %zu", width);
    if (stride == 0) return set_error(QD_E_ZERO_STRIDE, "attempt to divide by zero (fft.rs:86)");
    if (c->sharded()) {
        if (space != QD_SPACE_HOST) return set_error(QD_E_INVALID_ARG, "a sharded chain delivers into host buffers");
        uint64_t total = 0;
        QD_TRY(qd_freq_levels(c->shards[0], width, stride, levels, 0, 0, nullptr, space, &total));
        *total_out = total;
        if (first >= total || n == 0) return QD_OK;
        const uint64_t cnt = std::min(n, total - first);
        if (!vals) return set_error(QD_E_INVALID_ARG, "vals is null");
        const size_t parts = c->shards.size();
        std::vector<int> rc;
        std::vector<std::string> msg;
        run_on_shards(*c, [&](size_t i, qd_chain *s) {
            uint64_t a, b, t = 0;
            shard_range(cnt, parts, i, &a, &b);
            if (b == a) return static_cast<int>(QD_OK);
            return qd_freq_levels(s, width, stride, levels, first + a, b - a, vals + a, space, &t);
        }, rc, msg);
        return first_shard_error(rc, msg, nullptr);
    }
    std::lock_guard<std::mutex> lk(c->mu);
    uint64_t len = 0;
    QD_TRY(chain_len(*c, &len));
    const uint64_t total = (len - width) / stride; // u64 arithmetic as fft.rs:86 (wraps when len < width)
    *total_out = total;
    if (first >= total || n == 0) return QD_OK;
    const uint64_t cnt = std::min(n, total - first);
    if (!vals) return set_error(QD_E_INVALID_ARG, "vals is null");
    uint64_t good = 0;
    int bad = QD_OK;
    QD_TRY(first_bad_unit(*c, first * stride, stride, cnt, width, &good, &bad));
    if (bad != QD_OK) return bad; // .unwrap() at fft.rs:91: nothing is returned
    SinkArgs sink;
    sink.kind = SINK_LEVELS;
    sink.width = width;
    sink.space = space;
    sink.idx_out = vals;
    uint64_t produced = 0;
    return run_units(*c, first * stride, stride, nullptr, cnt, width, sink, &produced);
}

// rows [row0, row0 + n_rows) of take_fft; every row of the call is validated first, as the reference's `?` at
// ffts.rs:62 fails the whole call.  out receives the rows of the range only.
static int take_fft_rows(qd_chain *c, int has_slice, uint64_t start, uint64_t end, size_t width, int windowing,
                         size_t output_len, size_t row0, size_t n_rows, float *out, int space)
{
    std::lock_guard<std::mutex> lk(c->mu);
    uint64_t len = 0;
    QD_TRY(chain_len(*c, &len));
    const uint64_t start_sample = has_slice ? start : 0;
    const uint64_t end_sample = has_slice ? end : len - width; // ffts.rs:27-30
    if (!(end_sample > start_sample))
        return set_error(QD_E_SLICE, "Invalid slice: end (%llu) must be greater than start (%llu)",
                         (unsigned long long)end_sample, (unsigned long long)start_sample);
    if (!(end_sample < len))
        return set_error(QD_E_SLICE, "Slice end (%llu) exceeds sample length (%llu)", (unsigned long long)end_sample,
                         (unsigned long long)len);
    const uint64_t visible = end_sample - start_sample;
    if (!(visible > output_len))
        return set_error(QD_E_VISIBLE, "Visible samples (%llu) must be greater than output length (%zu)",
                         (unsigned long long)visible, output_len);
    if (output_len == 0) return QD_OK;
    if (!out) return set_error(QD_E_INVALID_ARG, "out is null");
    const double step = static_cast<double>(visible) / static_cast<double>(output_len); // ffts.rs:50
    std::vector<uint64_t> offs(output_len);
    for (size_t i = 0; i < output_len; i++) {
        offs[i] = start_sample + f64_as_u64(round(step * static_cast<double>(i))); // ffts.rs:60
        uint64_t v = 0;
        QD_TRY(chain_valid(*c, offs[i], width, &v));
        if (v != width) // read_exact_at(...)? at ffts.rs:62 propagates the Err
            return set_error(QD_E_SHORT_READ, "TODO: read-exact messed up: %zu (wanted) != %llu (read) at %llu", width,
                             (unsigned long long)v, (unsigned long long)offs[i]);
    }
    if (n_rows == 0) return QD_OK;
    SinkArgs sink;
    sink.kind = SINK_TAKE;
    sink.width = width;
    sink.windowed = windowing == 1;
    sink.space = space;
    sink.mag_out = out;
    uint64_t produced = 0;
    return run_units(*c, 0, 0, offs.data() + row0, n_rows, width, sink, &produced);
}

int qd_take_fft(qd_chain *c, int has_slice, uint64_t start, uint64_t end, size_t width, int windowing,
                size_t output_len, float *out, int space)
{
    if (!c) return set_error(QD_E_INVALID_ARG, "chain is null");
    // FftPlanner accepts any length (ffts.rs:25): powers of two run the radix-4 FFT, others a direct DFT
    if (width == 0 || width > 16384) return set_error(QD_E_FFT_WIDTH, "take_fft width %zu unsupported (1..16384)", width);
    if (!c->sharded()) return take_fft_rows(c, has_slice, start, end, width, windowing, output_len, 0, output_len, out, space);
    if (space != QD_SPACE_HOST) return set_error(QD_E_INVALID_ARG, "a sharded chain delivers into host buffers");
    const size_t parts = c->shards.size();
    std::vector<int> rc;
    std::vector<std::string> msg;
    run_on_shards(*c, [&](size_t i, qd_chain *s) {
        uint64_t a, b;
        shard_range(output_len, parts, i, &a, &b);
        return take_fft_rows(s, has_slice, start, end, width, windowing, output_len, a, b - a, out ? out + a * width : nullptr, space);
    }, rc, msg);
    return first_shard_error(rc, msg, nullptr);
}

// do_write's pull loop from sample offset `off`: at most n_reads reads of `chunk` samples, each starting where
// the previous one ended (lib.rs:200-204).  The chain's mutex is held by the caller.
static int write_from(qd_chain *c, size_t chunk, uint64_t off, uint64_t n_reads, qd_cf32 *out, uint64_t cap, int space,
                      uint64_t *n_out)
{
    *n_out = 0;
    if (c->sharded()) {
        // the run of FULL chunks goes to the devices in contiguous parts, each landing at its place of out[];
        // whatever follows (the ragged last read and the read that trips lib.rs:203) is one device's work
        if (space != QD_SPACE_HOST) return set_error(QD_E_INVALID_ARG, "a sharded chain delivers into host buffers");
        qd_chain *c0 = c->shards[0];
        uint64_t len = 0;
        QD_TRY(chain_len(*c0, &len));
        uint64_t good = 0;
        if (off < len && n_reads) {
            const uint64_t max_units = std::min<uint64_t>(n_reads, (len - off + chunk - 1) / chunk);
            int bad = QD_OK;
            QD_TRY(first_bad_unit(*c0, off, chunk, max_units, chunk, &good, &bad));
            good = std::min<uint64_t>(good, cap / chunk);
        }
        if (good) {
            if (!out) return set_error(QD_E_INVALID_ARG, "out is null");
            const size_t parts = c->shards.size();
            std::vector<uint64_t> got(parts, 0);
            std::vector<int> rc;
            std::vector<std::string> msg;
            run_on_shards(*c, [&](size_t i, qd_chain *s) {
                uint64_t a, b;
                shard_range(good, parts, i, &a, &b);
                if (b == a) return static_cast<int>(QD_OK);
                std::lock_guard<std::mutex> lk(s->mu);
                return write_from(s, chunk, off + a * chunk, b - a, out + a * chunk, (b - a) * chunk, space, &got[i]);
            }, rc, msg);
            QD_TRY(first_shard_error(rc, msg, nullptr));
            for (size_t i = 0; i < parts; i++) *n_out += got[i];
            if (*n_out != good * chunk) return set_error(QD_E_CUDA, "internal: sharded write produced %llu of %llu samples",
                                                         (unsigned long long)*n_out, (unsigned long long)(good * chunk));
        }
        if (good < n_reads && off + good * chunk < len) {
            uint64_t tail = 0;
            std::lock_guard<std::mutex> lk(c0->mu);
            const int rc = write_from(c0, chunk, off + good * chunk, n_reads - good, out ? out + good * chunk : nullptr,
                                      cap - good * chunk, space, &tail);
            *n_out += tail;
            return rc;
        }
        return QD_OK;
    }
    uint64_t len = 0;
    QD_TRY(chain_len(*c, &len));
    uint64_t done = 0, left = n_reads;
    // 1. the run of full chunks, as one batched launch sequence
    if (off < len && left) {
        uint64_t max_units = std::min<uint64_t>(left, (len - off + chunk - 1) / chunk);
        uint64_t good = 0;
        int bad = QD_OK;
        QD_TRY(first_bad_unit(*c, off, chunk, max_units, chunk, &good, &bad));
        good = std::min<uint64_t>(good, cap / chunk);
        if (good) {
            if (!out) return set_error(QD_E_INVALID_ARG, "out is null");
            SinkArgs sink;
            sink.kind = SINK_SAMPLES;
            sink.space = space;
            sink.samples_out = out;
            uint64_t produced = 0;
            QD_TRY(run_units(*c, off, chunk, nullptr, good, chunk, sink, &produced));
            done += produced;
            off += produced;
            left -= good;
            *n_out = done;
        }
    }
    // 2. the ragged tail, one read at a time exactly as lib.rs:200-204 advances `off`
    while (off < len && left) {
        uint64_t v = 0;
        QD_TRY(chain_valid(*c, off, chunk, &v));
        if (v == 0) // assert_ne!(0, read, "short read at offset {} of {}")
            return set_error(QD_E_WRITE_SHORT, "short read at offset %llu of %llu", (unsigned long long)off,
                             (unsigned long long)len);
        if (done + v > cap) return set_error(QD_E_INVALID_ARG, "output buffer too small (%llu samples)", (unsigned long long)cap);
        SinkArgs sink;
        sink.kind = SINK_SAMPLES;
        sink.space = space;
        sink.samples_out = out + done;
        uint64_t produced = 0;
        QD_TRY(run_units(*c, off, 0, nullptr, 1, chunk, sink, &produced));
        done += produced;
        off += produced;
        left--;
        *n_out = done;
    }
    return QD_OK;
}

int qd_write_cf32(qd_chain *c, size_t chunk, uint64_t first_chunk, uint64_t n_chunks, qd_cf32 *out, uint64_t cap,
                  int space, uint64_t *n_out)
{
    if (!c || !n_out) return set_error(QD_E_INVALID_ARG, "null argument");
    *n_out = 0;
    if (chunk == 0) return set_error(QD_E_INVALID_ARG, "chunk is 0");
    std::lock_guard<std::mutex> lk(c->mu);
    return write_from(c, chunk, first_chunk * chunk, n_chunks, out, cap, space, n_out);
}

int qd_write_file(qd_chain *c, const char *prefix, int overwrite, char *name_out, size_t name_cap)
{
    if (!c || !prefix) return set_error(QD_E_INVALID_ARG, "null argument");
    if (strcmp(prefix, "-") == 0) return set_error(QD_E_UNIMPLEMENTED, "not implemented"); // lib.rs:179-181
    char name[4096];
    snprintf(name, sizeof name, "%s.sr%llu.cf32", prefix, (unsigned long long)chain_rate(c->sharded() ? *c->shards[0] : *c)); // lib.rs:194
    if (name_out && name_cap) snprintf(name_out, name_cap, "%s", name);
    const int fd = open(name, O_WRONLY | O_CREAT | (overwrite ? 0 : O_EXCL), 0666); // lib.rs:186-192
    if (fd < 0) return set_error(errno == EEXIST ? QD_E_EXISTS : QD_E_IO, "%s: %s", name, strerror(errno));
    FILE *f = fdopen(fd, "wb");
    if (!f) {
        const int e = errno;
        close(fd);
        return set_error(QD_E_IO, "%s: %s", name, strerror(e));
    }
    const size_t chunk = 0x1000; // lib.rs:201
    const uint64_t batch = 512 * (c->sharded() ? c->shards.size() : 1); // reads per device round trip
    std::vector<qd_cf32> host(chunk * batch);
    int rc = QD_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    uint64_t len = 0;
    rc = chain_len(c->sharded() ? *c->shards[0] : *c, &len);
    // `while off < len` (lib.rs:200): every batch continues at the sample where the previous one ended, so a
    // ragged short read that happens to be the last read of a batch is still followed by the read that returns
    // 0 samples and trips the reference's assert_ne! (lib.rs:203 -> QD_E_WRITE_SHORT)
    for (uint64_t off = 0; rc == QD_OK && off < len;) {
        uint64_t n = 0;
        rc = write_from(c, chunk, off, batch, host.data(), host.size(), QD_SPACE_HOST, &n);
        if (n && fwrite(host.data(), sizeof(qd_cf32), n, f) != n) { // LE f32 re, im (lib.rs:206-209)
            rc = set_error(QD_E_IO, "%s: %s", name, strerror(errno));
            break;
        }
        off += n;
        if (rc == QD_OK && n == 0) break; // cannot happen: a read of 0 samples is QD_E_WRITE_SHORT
    }
    if (fclose(f) != 0 && rc == QD_OK) rc = set_error(QD_E_IO, "%s: %s", name, strerror(errno));
    return rc;
}

int qd_shard_plan(const qd_source *src, const qd_stage *stages, size_t n_stages, int sink_kind, uint64_t unit_len,
                  uint64_t stride, uint32_t n_shards, uint32_t shard, qd_shard *out)
{
    if (!src || !out || n_shards == 0 || shard >= n_shards || unit_len == 0)
        return set_error(QD_E_INVALID_ARG, "qd_shard_plan: bad arguments");
    Chain c;
    qd_source s = *src;
    if (s.kind == QD_SRC_HOST_MEM || s.kind == QD_SRC_DEVICE_MEM) { // geometry only: data may be absent
        s.kind = QD_SRC_HOST_MEM;
        if (!s.data) s.data = "";
    }
    QD_TRY(describe(&s, stages, n_stages, c));
    uint64_t len = 0;
    QD_TRY(chain_len(c, &len));
    uint64_t total, step;
    if (sink_kind == 0) { // do_write: reads at 0, chunk, 2*chunk, ... while off < len
        total = (len + unit_len - 1) / unit_len;
        step = unit_len;
    } else {
        if (stride == 0) return set_error(QD_E_ZERO_STRIDE, "stride is 0");
        if (sink_kind == 1) total = len > unit_len ? (len - unit_len + stride - 1) / stride : 0; // fft.rs:27-28
        else total = len >= unit_len ? (len - unit_len) / stride : 0;                             // fft.rs:86
        step = stride;
    }
    const uint64_t a = total * shard / n_shards, b = total * (shard + 1) / n_shards;
    out->first_unit = a;
    out->n_units = b - a;
    out->first_sample = 0;
    out->n_samples = 0;
    if (b > a) {
        uint64_t lo, hi, lo2, hi2;
        chain_source_span(c, a * step, unit_len, &lo, &hi);
        chain_source_span(c, (b - 1) * step, unit_len, &lo2, &hi2);
        out->first_sample = lo;
        out->n_samples = std::max(hi, hi2) - lo;
    }
    return QD_OK;
}

int qd_synth_fill(const qd_synth *p, int format, uint64_t first_sample, uint64_t n_samples, void *device_out,
                  int device, void *cuda_stream)
{
    return synth_fill(p, format, first_sample, n_samples, device_out, device, static_cast<cudaStream_t>(cuda_stream));
}

} // extern "C"
