/*
 * oracle/quadrs_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of quadrs's streaming IQ DSP chain.  See quadrs_oracle.h for
 * the parity status ("FFT parity unpinned").  Every function cites the
 * reference file:line (relative to /root/reference) whose behaviour it
 * restates.  The structure deliberately keeps the reference's lazy pull graph
 * (per-call allocation, full-rate convolve, per-sample f64 sin/cos) because
 * this file is also the timed CPU baseline.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  All
 * f32/f64 operations below are individually rounded IEEE operations, as in
 * Rust (which never contracts to FMA).  libm calls (sinf, cosf, sin, cos,
 * hypotf) are the glibc ones Rust's std calls on linux-gnu.
 */
#define _GNU_SOURCE
#include "quadrs_oracle.h"

#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <pthread.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------ */
/* panic emulation: Rust assert!/expect/index panics -> status codes   */
/* ------------------------------------------------------------------ */

static __thread jmp_buf *tl_jmp;
static __thread int tl_code;
static __thread char tl_msg[320];
static int g_kept_only = 0;

const char *qo_last_error(void) { return tl_msg; }
void qo_set_kept_only_convolve(int on) { g_kept_only = on; }

static void set_msg(const char *fmt, va_list ap) { vsnprintf(tl_msg, sizeof tl_msg, fmt, ap); }

__attribute__((noreturn, format(printf, 2, 3))) static void qo_panic(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    set_msg(fmt, ap);
    va_end(ap);
    tl_code = code;
    if (!tl_jmp) {
        fprintf(stderr, "quadrs_oracle: panic outside API call: %s\n", tl_msg);
        abort();
    }
    longjmp(*tl_jmp, 1);
}

__attribute__((format(printf, 2, 3))) static int qo_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    set_msg(fmt, ap);
    va_end(ap);
    return code;
}

#define QO_ENTER()                                                                                                     \
    jmp_buf jb_;                                                                                                       \
    jmp_buf *prev_jmp_ = tl_jmp;                                                                                       \
    tl_jmp = &jb_;                                                                                                     \
    if (setjmp(jb_)) {                                                                                                 \
        tl_jmp = prev_jmp_;                                                                                            \
        return tl_code;                                                                                                \
    }
#define QO_LEAVE(rc)                                                                                                   \
    do {                                                                                                               \
        tl_jmp = prev_jmp_;                                                                                            \
        return (rc);                                                                                                   \
    } while (0)

static void *xcalloc(size_t n, size_t sz)
{
    void *p = calloc(n ? n : 1, sz);
    if (!p) qo_panic(QO_E_NOMEM, "out of memory (%zu x %zu)", n, sz);
    return p;
}

/* ------------------------------------------------------------------ */
/* graph nodes                                                         */
/* ------------------------------------------------------------------ */

enum { K_FILE = 0, K_GEN = 1, K_SHIFT = 2, K_LOWPASS = 3 };

struct qo_samples {
    int kind;
    qo_samples *inner; /* owned, as Shift<S>{inner} / LowPass<S>{inner} own by value (shift.rs:7-11, filter.rs:14-19) */
    /* SampleFile, samples.rs:44-49 */
    const uint8_t *mem;
    uint64_t mem_base_bytes; /* window source (tests only): mem[0] is byte mem_base_bytes of the capture */
    uint64_t mem_bytes;
    int fd;
    uint64_t file_len;
    int format;
    uint64_t sample_rate;
    /* Shift, shift.rs:7-11 */
    double ratio;
    /* LowPass, filter.rs:14-19 */
    float *filter;
    size_t filter_len;
    uint64_t decimate;
    /* Gen, gen.rs:10-14 */
    double seconds;
    int64_t *cos_hz;
    size_t n_cos;
};

/* FileFormat::type_bytes / pair_bytes, lib.rs:217-229 */
static uint64_t pair_bytes(int format)
{
    switch (format) {
    case QO_CF32: return 8;
    case QO_CS8:
    case QO_CU8: return 2;
    case QO_CS16: return 4;
    }
    return 0;
}

/* FileFormat::to_f32, lib.rs:241-255.  Little-endian host assumed (x86-64). */
static inline float to_f32(int format, const uint8_t *b)
{
    switch (format) {
    case QO_CF32: {
        float f;
        memcpy(&f, b, 4); /* LittleEndian::read_f32: bit copy */
        return f;
    }
    case QO_CS8: return (float)(int8_t)b[0] / 127.0f;
    case QO_CU8: return (float)b[0] / 255.0f - (255.0f / 2.0f);
    case QO_CS16: {
        int16_t v;
        memcpy(&v, b, 2);
        return (float)v / 65535.0f - (65535.0f / 2.0f);
    }
    }
    return 0.0f;
}

/* FileFormat::to_cf32, lib.rs:231-238: I first, Q second */
static inline qo_cf32 to_cf32(int format, const uint8_t *b)
{
    size_t tb = (size_t)pair_bytes(format) / 2;
    qo_cf32 c = {to_f32(format, b), to_f32(format, b + tb)};
    return c;
}

void qo_decode(int format, const uint8_t *in, size_t n, qo_cf32 *out)
{
    size_t pb = (size_t)pair_bytes(format);
    for (size_t i = 0; i < n; i++) out[i] = to_cf32(format, in + i * pb);
}

qo_samples *qo_from_mem(const uint8_t *data, uint64_t n_bytes, int format, uint64_t sample_rate)
{
    if (!pair_bytes(format)) return NULL;
    qo_samples *s = calloc(1, sizeof *s);
    if (!s) return NULL;
    s->kind = K_FILE;
    s->mem = data;
    s->mem_bytes = n_bytes;
    s->fd = -1;
    s->file_len = n_bytes;
    s->format = format;
    s->sample_rate = sample_rate;
    return s;
}

/* Test helper, not in the reference: a capture of total_samples of which only the window
 * [base_sample, base_sample + n_bytes/pair) is backed by memory.  Lets the oracle evaluate reads at
 * absolute offsets near 2^33..2^34 without materialising 32..128 GiB. */
qo_samples *qo_from_mem_window(const uint8_t *data, uint64_t n_bytes, int format, uint64_t sample_rate,
                               uint64_t base_sample, uint64_t total_samples)
{
    qo_samples *s = qo_from_mem(data, n_bytes, format, sample_rate);
    if (!s) return NULL;
    s->mem_base_bytes = base_sample * pair_bytes(format);
    s->file_len = total_samples * pair_bytes(format);
    return s;
}

/* SampleFile::new, samples.rs:51-61: length = seek(End) */
qo_samples *qo_from_file(const char *path, int format, uint64_t sample_rate)
{
    if (!pair_bytes(format)) return NULL;
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        qo_err(QO_E_IO, "open %s: %s", path, strerror(errno));
        return NULL;
    }
    off_t end = lseek(fd, 0, SEEK_END);
    if (end < 0) {
        close(fd);
        return NULL;
    }
    qo_samples *s = calloc(1, sizeof *s);
    if (!s) {
        close(fd);
        return NULL;
    }
    s->kind = K_FILE;
    s->fd = fd;
    s->file_len = (uint64_t)end;
    s->format = format;
    s->sample_rate = sample_rate;
    return s;
}

/* Gen::new, gen.rs:17-27 */
int qo_gen(const int64_t *cos_hz, size_t n_cos, uint64_t sample_rate, double seconds, qo_samples **out)
{
    if (n_cos == 0) return qo_err(QO_E_GEN_ARGS, "cos cannot be empty");
    if (sample_rate == 0) return qo_err(QO_E_GEN_ARGS, "sample rate may not be zero");
    if (!(seconds > 0.0)) return qo_err(QO_E_GEN_ARGS, "seconds may not be <= 0");
    qo_samples *s = calloc(1, sizeof *s);
    if (!s) return QO_E_NOMEM;
    s->kind = K_GEN;
    s->fd = -1;
    s->sample_rate = sample_rate;
    s->seconds = seconds;
    s->cos_hz = malloc(n_cos * sizeof(int64_t));
    memcpy(s->cos_hz, cos_hz, n_cos * sizeof(int64_t));
    s->n_cos = n_cos;
    *out = s;
    return QO_OK;
}

/* TAU, lib.rs:12,23: std::f64::consts::PI * 2. */
static const double QO_TAU = M_PI * 2.0;

/* shift.rs:28: TAU * (frequency as f64) / (sample_rate as f64), left-assoc */
double qo_shift_ratio(int64_t frequency, uint64_t sample_rate)
{
    return QO_TAU * (double)frequency / (double)sample_rate;
}

/* Operation::Shift arm lib.rs:102-106 + Shift::new shift.rs:19-31.  The sample
 * rate handed to Shift::new is the inner stage's. */
int qo_shift(qo_samples *inner, int64_t frequency, qo_samples **out)
{
    if (!inner) return qo_err(QO_E_INVALID_ARG, "shift requires an input");
    uint64_t sr = qo_sample_rate(inner);
    int64_t a = frequency < 0 ? -frequency : frequency; /* i64::abs */
    if (!(a < (int64_t)(sr / 2))) return qo_err(QO_E_SHIFT_NYQUIST, "frequency must be under half the sample rate");
    if (!(sr > 0)) return qo_err(QO_E_ZERO_RATE, "assertion failed: sample_rate > 0");
    qo_samples *s = calloc(1, sizeof *s);
    if (!s) return QO_E_NOMEM;
    s->kind = K_SHIFT;
    s->fd = -1;
    s->inner = inner;
    s->ratio = qo_shift_ratio(frequency, sr);
    s->sample_rate = sr;
    *out = s;
    return QO_OK;
}

/* lowpass_filter, filter.rs:86-105, with cutoff from filter.rs:29-31,126-128.
 * f32 throughout; PI is std::f32::consts::PI. */
static void lowpass_filter(float cutoff, size_t size, float *filter)
{
    const float PI = 3.14159274101257324219f; /* f32::consts::PI */
    for (size_t i = 0; i < size; i++) {
        /* sinc(2.0 * cutoff * (i as f32 - (size as f32 - 1.0) / 2.0)) */
        float x = 2.0f * cutoff * ((float)i - ((float)size - 1.0f) / 2.0f);
        float xp = x * PI;
        float wave = sinf(xp) / xp;
        /* 0.42 - 0.5 * cos(2.0*PI*i/(size-1)) + 0.08 * cos(4.0*PI*i/(size-1)) */
        float a1 = 2.0f * PI * (float)i / ((float)size - 1.0f);
        float a2 = 4.0f * PI * (float)i / ((float)size - 1.0f);
        float window = 0.42f - 0.5f * cosf(a1) + 0.08f * cosf(a2);
        filter[i] = wave * window;
    }
    /* Normalize: sequential f32 sum (Iterator::sum), then divide */
    float sum = 0.0f;
    for (size_t i = 0; i < size; i++) sum = sum + filter[i];
    for (size_t i = 0; i < size; i++) filter[i] = filter[i] / sum;
}

int qo_taps(uint64_t frequency, uint64_t sample_rate, size_t size, float *out)
{
    /* cutoff_from_frequency(frequency as f64, sr) -> f64, then `cutoff as f32` (filter.rs:29-31) */
    double cutoff = (double)frequency / (double)sample_rate;
    lowpass_filter((float)cutoff, size, out);
    return QO_OK;
}

/* Operation::LowPass arm lib.rs:107-121 + LowPass::new filter.rs:21-40 */
int qo_lowpass(qo_samples *inner, uint64_t frequency, uint64_t decimate, size_t size, qo_samples **out)
{
    if (!inner) return qo_err(QO_E_INVALID_ARG, "lowpass requires an input");
    qo_samples *s = calloc(1, sizeof *s);
    if (!s) return QO_E_NOMEM;
    s->kind = K_LOWPASS;
    s->fd = -1;
    s->inner = inner;
    s->sample_rate = qo_sample_rate(inner); /* original_sample_rate */
    s->decimate = decimate;
    s->filter_len = size;
    s->filter = malloc((size ? size : 1) * sizeof(float));
    qo_taps(frequency, s->sample_rate, size, s->filter);
    *out = s;
    return QO_OK;
}

void qo_free(qo_samples *s)
{
    if (!s) return;
    qo_free(s->inner);
    if (s->kind == K_FILE && s->fd >= 0) close(s->fd);
    free(s->filter);
    free(s->cos_hz);
    free(s);
}

/* ------------------------------------------------------------------ */
/* trait Samples                                                       */
/* ------------------------------------------------------------------ */

static uint64_t f64_as_u64(double v) /* Rust `as u64`: saturating, NaN -> 0 */
{
    if (!(v > 0.0)) return 0;
    if (v >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)v;
}

static uint64_t s_len(const qo_samples *s)
{
    switch (s->kind) {
    case K_FILE: return s->file_len / pair_bytes(s->format);              /* samples.rs:64-66 */
    case K_GEN: return f64_as_u64(s->seconds * (double)s->sample_rate);  /* gen.rs:31-33 */
    case K_SHIFT: return s_len(s->inner);                                 /* shift.rs:38-40 */
    case K_LOWPASS: {                                                     /* filter.rs:45-48 */
        uint64_t il = s_len(s->inner);
        if (!(il >= (uint64_t)s->filter_len))
            qo_panic(QO_E_SHORT_INPUT, "assertion failed: self.inner.len() >= self.filter.len() as u64");
        if (s->decimate == 0) qo_panic(QO_E_INVALID_ARG, "attempt to divide by zero");
        return 1 + (il - (uint64_t)s->filter_len) / s->decimate;
    }
    }
    return 0;
}

uint64_t qo_sample_rate(const qo_samples *s)
{
    switch (s->kind) {
    case K_LOWPASS: return s->decimate ? s->sample_rate / s->decimate : 0; /* filter.rs:50-52 */
    default: return s->sample_rate;
    }
}

/* complex_convolve, filter.rs:107-124 (literal).  Returns valid-1+L/2 items. */
static qo_cf32 *complex_convolve(const float *filter, size_t flen, const qo_cf32 *input, size_t ilen, size_t *olen)
{
    long h_len = (long)(flen / 2);
    long lo = -((long)flen / 2), hi = (long)ilen - 1;
    size_t cap = ilen + flen / 2 + 1;
    qo_cf32 *output = xcalloc(cap, sizeof(qo_cf32));
    size_t count = 0;
    for (long i = lo; i < hi; i++) {
        size_t output_idx = (size_t)(i + h_len);
        qo_cf32 acc = {0.0f, 0.0f}; /* output.push(Complex::zero()) */
        for (long j = 0; j < (long)flen; j++) {
            long input_idx = i + j;
            if (input_idx < 0 || input_idx >= (long)ilen) continue;
            /* output[idx] += input[input_idx] * filter[j]: Complex*f32 then Complex+= */
            float pr = input[input_idx].re * filter[j];
            float pi = input[input_idx].im * filter[j];
            acc.re = acc.re + pr;
            acc.im = acc.im + pi;
        }
        output[output_idx] = acc;
        count++;
    }
    *olen = count;
    return output;
}

static size_t s_read_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n);

/* SampleFile::read_at, samples.rs:72-93 */
static size_t file_read_at(const qo_samples *s, uint64_t off, qo_cf32 *into, size_t n)
{
    uint64_t pb = pair_bytes(s->format);
    if (!(off < s_len(s))) qo_panic(QO_E_OFFSET_EOF, "assertion failed: off < self.len()");
    size_t wanted_bytes;
    if (__builtin_mul_overflow((size_t)pb, n, &wanted_bytes)) qo_panic(QO_E_INVALID_ARG, "buf too big");
    size_t bytes;
    uint8_t *tmp = NULL;
    const uint8_t *src;
    if (s->mem) {
        uint64_t avail = s->file_len - off * pb;
        bytes = wanted_bytes < avail ? wanted_bytes : (size_t)avail;
        /* the reference allocates and fills a Vec<u8> per call (samples.rs:79-83) */
        tmp = xcalloc(wanted_bytes, 1);
        if (off * pb < s->mem_base_bytes || off * pb + bytes > s->mem_base_bytes + s->mem_bytes) {
            free(tmp);
            qo_panic(QO_E_INVALID_ARG, "window source: bytes [%llu, +%zu) are not backed", (unsigned long long)(off * pb),
                     bytes);
        }
        memcpy(tmp, s->mem + (off * pb - s->mem_base_bytes), bytes);
        src = tmp;
    } else {
        tmp = xcalloc(wanted_bytes, 1);
        ssize_t r = pread(s->fd, tmp, wanted_bytes, (off_t)(off * pb));
        if (r < 0) {
            free(tmp);
            qo_panic(QO_E_IO, "read: %s", strerror(errno));
        }
        bytes = (size_t)r;
        src = tmp;
    }
    bytes -= bytes % (size_t)pb;
    size_t cnt = bytes / (size_t)pb;
    for (size_t i = 0; i < cnt; i++) into[i] = to_cf32(s->format, src + i * (size_t)pb);
    free(tmp);
    return cnt;
}

/* Gen::read_at, gen.rs:35-47: ignores len(), always fills the buffer */
static size_t gen_read_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        double base = (double)(off + (uint64_t)i) * QO_TAU / (double)s->sample_rate;
        qo_cf32 val = {0.0f, 0.0f};
        for (size_t t = 0; t < s->n_cos; t++) {
            double f = (double)s->cos_hz[t] * base;
            val.re = val.re + (float)cos(f);
            val.im = val.im + (float)sin(f);
        }
        buf[i] = val;
    }
    return n;
}

/* Shift::read_at, shift.rs:46-54; Complex *= per num-complex 0.4.6 MulAssign:
 * re' = re*c - im*s ; im' = im*c + re*s, each op rounded. */
static size_t shift_read_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n)
{
    size_t valid = s_read_at(s->inner, off, buf, n);
    for (size_t i = 0; i < valid; i++) {
        double place = (double)(off + (uint64_t)i) * s->ratio;
        float c = (float)cos(place), sn = (float)sin(place);
        float a = buf[i].re, b = buf[i].im;
        float re = a * c - b * sn;
        float im = b * c + a * sn;
        buf[i].re = re;
        buf[i].im = im;
    }
    return valid;
}

/* LowPass::read_at, filter.rs:54-83 */
static size_t lowpass_read_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n)
{
    size_t L = s->filter_len;
    size_t D = (size_t)s->decimate;
    size_t underlying = n * D + L;
    qo_cf32 *raw = xcalloc(underlying, sizeof(qo_cf32));
    size_t valid = s_read_at(s->inner, off * s->decimate, raw, underlying);
    size_t output_samples;
    if (!g_kept_only) {
        size_t clen;
        qo_cf32 *conv = complex_convolve(s->filter, L, raw, valid, &clen);
        if (L / 2 - 1 + valid != clen) { /* assert_eq!, filter.rs:74 */
            free(conv);
            free(raw);
            qo_panic(QO_E_INVALID_ARG, "assertion failed: filter.len()/2 - 1 + valid == convoluted.len()");
        }
        if (valid < L || D == 0) { /* filter.rs:76: usize underflow -> panic (debug) / OOB index (release) */
            free(conv);
            free(raw);
            qo_panic(QO_E_SHORT_INPUT, "attempt to subtract with overflow (valid %zu < filter %zu)", valid, L);
        }
        output_samples = (valid - L) / D;
        for (size_t i = 0; i < output_samples; i++) buf[i] = conv[L + i * D]; /* filter.rs:78-80 */
        free(conv);
    } else {
        /* Same values, kept outputs only: convoluted[L + k*D] is loop index
         * i = L + k*D - L/2 of filter.rs:111, taps ascending j, input_idx < valid. */
        if (valid < L || D == 0) {
            free(raw);
            qo_panic(QO_E_SHORT_INPUT, "attempt to subtract with overflow (valid %zu < filter %zu)", valid, L);
        }
        output_samples = (valid - L) / D;
        size_t h = L / 2;
        for (size_t k = 0; k < output_samples; k++) {
            size_t i0 = L + k * D - h;
            qo_cf32 acc = {0.0f, 0.0f};
            for (size_t j = 0; j < L; j++) {
                size_t idx = i0 + j;
                if (idx >= valid) continue;
                float pr = raw[idx].re * s->filter[j];
                float pi = raw[idx].im * s->filter[j];
                acc.re = acc.re + pr;
                acc.im = acc.im + pi;
            }
            buf[k] = acc;
        }
    }
    free(raw);
    return output_samples;
}

static size_t s_read_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n)
{
    switch (s->kind) {
    case K_FILE: return file_read_at(s, off, buf, n);
    case K_GEN: return gen_read_at(s, off, buf, n);
    case K_SHIFT: return shift_read_at(s, off, buf, n);
    case K_LOWPASS: return lowpass_read_at(s, off, buf, n);
    }
    return 0;
}

/* Samples::read_exact_at, samples.rs:17-27: Err (not panic) on short read */
static int s_read_exact_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n)
{
    size_t got = s_read_at(s, off, buf, n);
    if (got != n)
        return qo_err(QO_E_SHORT_READ, "TODO: read-exact messed up: %zu (wanted) != %zu (read) at %llu", n, got,
                      (unsigned long long)off);
    return QO_OK;
}

int qo_len(const qo_samples *s, uint64_t *out)
{
    QO_ENTER();
    *out = s_len(s);
    QO_LEAVE(QO_OK);
}

int qo_read_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n, size_t *got)
{
    QO_ENTER();
    *got = s_read_at(s, off, buf, n);
    QO_LEAVE(QO_OK);
}

int qo_read_exact_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n)
{
    QO_ENTER();
    int rc = s_read_exact_at(s, off, buf, n);
    QO_LEAVE(rc);
}

/* ------------------------------------------------------------------ */
/* FFT: our own definition (rustfft 6.4.0 is not in /root/reference)    */
/* ------------------------------------------------------------------ */
/*
 * Forward, unnormalised, kernel e^{-2 pi i jk/N}, N a power of two.
 * Radix-4 decimation in time; a radix-2 butterfly is the innermost stage when
 * log2 N is odd.  Twiddle w(N, j) = (f32 cos a, f32 sin a), a = (-2 pi / N) * j
 * in f64 (rustfft builds its twiddles in f64 and casts [recall]).  A twiddle
 * with index 0 is not multiplied.  Complex multiply is num-complex's
 * (ar*br - ai*bi, ar*bi + ai*br), each op rounded, no FMA.  The GPU kernels
 * evaluate exactly this DAG, so GPU == oracle bit for bit by construction.
 */
static void fft_twiddles(size_t N, qo_cf32 *T) /* T[j] = w(N, j), j < N */
{
    double constant = -2.0 * M_PI / (double)N;
    for (size_t j = 0; j < N; j++) {
        double angle = constant * (double)j;
        T[j].re = (float)cos(angle);
        T[j].im = (float)sin(angle);
    }
}

static inline qo_cf32 cmul(qo_cf32 a, qo_cf32 b)
{
    qo_cf32 r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
    return r;
}

static void fft_rec(const qo_cf32 *in, size_t stride, size_t n, qo_cf32 *out, const qo_cf32 *T, size_t N)
{
    if (n == 1) {
        out[0] = in[0];
        return;
    }
    if (n == 2) {
        qo_cf32 a = in[0], b = in[stride];
        out[0].re = a.re + b.re;
        out[0].im = a.im + b.im;
        out[1].re = a.re - b.re;
        out[1].im = a.im - b.im;
        return;
    }
    size_t q = n / 4;
    for (size_t r = 0; r < 4; r++) fft_rec(in + r * stride, stride * 4, q, out + r * q, T, N);
    size_t scale = N / n; /* w(n, j) == T[j * N/n] exactly (power-of-two scaling of the angle) */
    for (size_t k = 0; k < q; k++) {
        qo_cf32 t0 = out[k], t1 = out[q + k], t2 = out[2 * q + k], t3 = out[3 * q + k];
        if (k != 0) {
            t1 = cmul(t1, T[k * scale]);
            t2 = cmul(t2, T[2 * k * scale]);
            t3 = cmul(t3, T[3 * k * scale]);
        }
        qo_cf32 s0 = {t0.re + t2.re, t0.im + t2.im};
        qo_cf32 s1 = {t0.re - t2.re, t0.im - t2.im};
        qo_cf32 s2 = {t1.re + t3.re, t1.im + t3.im};
        qo_cf32 s3 = {t1.re - t3.re, t1.im - t3.im};
        out[k].re = s0.re + s2.re;
        out[k].im = s0.im + s2.im;
        out[q + k].re = s1.re + s3.im; /* s1 - i*s3 */
        out[q + k].im = s1.im - s3.re;
        out[2 * q + k].re = s0.re - s2.re;
        out[2 * q + k].im = s0.im - s2.im;
        out[3 * q + k].re = s1.re - s3.im; /* s1 + i*s3 */
        out[3 * q + k].im = s1.im + s3.re;
    }
}

typedef struct {
    size_t n;
    qo_cf32 *T;
    qo_cf32 *scratch;
} fft_plan;

static int is_pow2(size_t n) { return n && !(n & (n - 1)); }

/* FftPlanner::plan_fft_forward (ffts.rs:25) accepts ANY length; rustfft's mixed-radix / Rader / Bluestein
 * code is not in the reference tree.  Our definition for a length that is not a power of two is the direct
 * DFT  X[k] = sum_j x[j] * w(N, (j*k) mod N)  with the f32 twiddles above, products and the running sum in
 * f64 (an f32 x f32 product is exact in f64, so every term is rounded once, on addition), ascending j,
 * rounded to f32 at the end.  The GPU kernel gk_dft evaluates exactly this. */
static void plan_init_any(fft_plan *p, size_t n)
{
    if (n == 0) qo_panic(QO_E_FFT_WIDTH, "FFT length 0");
    p->n = n;
    p->T = xcalloc(n, sizeof(qo_cf32));
    p->scratch = xcalloc(n, sizeof(qo_cf32));
    fft_twiddles(n, p->T);
}

static void plan_process_any(fft_plan *p, qo_cf32 *buf)
{
    if (is_pow2(p->n)) {
        fft_rec(buf, 1, p->n, p->scratch, p->T, p->n);
    } else {
        const size_t n = p->n;
        for (size_t k = 0; k < n; k++) {
            double re = 0.0, im = 0.0;
            size_t m = 0; /* (j*k) mod n */
            for (size_t j = 0; j < n; j++) {
                const double xr = buf[j].re, xi = buf[j].im, wr = p->T[m].re, wi = p->T[m].im;
                re = re + xr * wr;
                re = re - xi * wi;
                im = im + xr * wi;
                im = im + xi * wr;
                m += k;
                if (m >= n) m -= n;
            }
            p->scratch[k].re = (float)re;
            p->scratch[k].im = (float)im;
        }
    }
    memcpy(buf, p->scratch, p->n * sizeof(qo_cf32));
}

static void plan_init(fft_plan *p, size_t n)
{
    if (!is_pow2(n)) qo_panic(QO_E_FFT_WIDTH, "Radix4 algorithm requires a power-of-two input size. Got %zu", n);
    p->n = n;
    p->T = xcalloc(n, sizeof(qo_cf32));
    p->scratch = xcalloc(n, sizeof(qo_cf32));
    fft_twiddles(n, p->T);
}

static void plan_free(fft_plan *p)
{
    free(p->T);
    free(p->scratch);
}

static void plan_process(fft_plan *p, qo_cf32 *buf)
{
    fft_rec(buf, 1, p->n, p->scratch, p->T, p->n);
    memcpy(buf, p->scratch, p->n * sizeof(qo_cf32));
}

int qo_fft(qo_cf32 *buf, size_t n)
{
    QO_ENTER();
    fft_plan p;
    plan_init(&p, n);
    plan_process(&p, buf);
    plan_free(&p);
    QO_LEAVE(QO_OK);
}

void qo_dft_c128(const qo_cf32 *in, size_t n, double *out)
{
    for (size_t k = 0; k < n; k++) {
        long double sr = 0, si = 0;
        for (size_t j = 0; j < n; j++) {
            size_t m = (j * k) % n;
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n;
            long double c = cosl(a), s = sinl(a);
            sr += (long double)in[j].re * c - (long double)in[j].im * s;
            si += (long double)in[j].re * s + (long double)in[j].im * c;
        }
        out[2 * k] = (double)sr;
        out[2 * k + 1] = (double)si;
    }
}

/* ------------------------------------------------------------------ */
/* sinks                                                               */
/* ------------------------------------------------------------------ */

/* fft.rs:45,53-60.  `as usize` saturates; graph has 7 entries. */
int qo_glyph_index(float norm, float min, float max)
{
    if (norm < min) return 0;
    if (norm >= max) return 8;
    float distinction = (max - min) / 7.0f;
    float q = (norm - min) / distinction;
    size_t idx;
    if (!(q > 0.0f)) idx = 0; /* NaN and negatives saturate to 0 */
    else if (q >= 1.8e19f) idx = SIZE_MAX;
    else idx = (size_t)q;
    if (idx >= 7) return -1; /* graph[idx] panics */
    return 1 + (int)idx;
}

static const char *const GLYPHS[9] = {" ", "▁", "▂", "▃", "▄", "▅", "▆", "▇", "█"};

/* fft.rs:63: println!("│{}│", buf) without the newline */
size_t qo_format_row(const uint8_t *idx, size_t width, char *out)
{
    size_t o = 0;
    memcpy(out + o, "│", 3);
    o += 3;
    for (size_t b = 0; b < width; b++) {
        const char *g = GLYPHS[idx[b] <= 8 ? idx[b] : 0];
        size_t l = strlen(g);
        memcpy(out + o, g, l);
        o += l;
    }
    memcpy(out + o, "│", 3);
    o += 3;
    return o;
}

int qo_spark_rows(const qo_samples *s, size_t width, uint64_t stride, uint64_t *rows)
{
    QO_ENTER();
    if (stride == 0) QO_LEAVE(qo_err(QO_E_ZERO_STRIDE, "stride 0 never terminates (fft.rs:65)"));
    uint64_t len = s_len(s);
    if (len <= (uint64_t)width) {
        /* len - width wraps (release) : the first read_exact_at fails; no rows */
        *rows = 0;
        QO_LEAVE(QO_OK);
    }
    uint64_t span = len - (uint64_t)width;
    *rows = (span + stride - 1) / stride;
    QO_LEAVE(QO_OK);
}

/* spark_fft, fft.rs:12-69 */
int qo_spark_fft(qo_samples *s, size_t width, uint64_t stride, int has_min, float min_in, int has_max, float max_in,
                 uint64_t first_row, uint64_t max_rows, uint8_t *idx_out, float *mag_out, uint64_t *rows_out)
{
    QO_ENTER();
    *rows_out = 0;
    float min = has_min ? min_in : 0.08f; /* fft.rs:22-23 */
    float max = has_max ? max_in : 1.0f;
    if (stride == 0) QO_LEAVE(qo_err(QO_E_ZERO_STRIDE, "stride 0 never terminates (fft.rs:65)"));
    fft_plan plan;
    plan_init(&plan, width); /* Radix4::new, fft.rs:25 */
    qo_cf32 *inp = xcalloc(width, sizeof(qo_cf32));
    uint64_t limit = s_len(s) - (uint64_t)width; /* u64 arithmetic as in fft.rs:28 (wraps in release) */
    uint64_t produced = 0;
    /* row r of the reference loop starts at i = r*stride; jump straight to first_row */
    if (first_row && first_row > UINT64_MAX / stride) {
        free(inp);
        plan_free(&plan);
        QO_LEAVE(QO_OK);
    }
    for (uint64_t i = first_row * stride; i < limit; i += stride) {
        if (produced >= max_rows) break;
        memset(inp, 0, width * sizeof(qo_cf32));
        int rc = s_read_exact_at(s, i, inp, width); /* fft.rs:29-30 */
        if (rc != QO_OK) {
            free(inp);
            plan_free(&plan);
            *rows_out = produced;
            QO_LEAVE(rc);
        }
        plan_process(&plan, inp); /* fft.rs:32 */
        /* iter().skip(w/2).chain(iter().take(w/2)), fft.rs:48-52 */
        for (size_t b = 0; b < width; b++) {
            size_t src = b < width - width / 2 ? b + width / 2 : b - (width - width / 2);
            float norm = hypotf(inp[src].re, inp[src].im); /* Complex::norm = re.hypot(im) */
            int g = qo_glyph_index(norm, min, max);
            if (g < 0) {
                free(inp);
                plan_free(&plan);
                *rows_out = produced;
                QO_LEAVE(qo_err(QO_E_GLYPH_RANGE, "index out of bounds: the len is 7 but the index is 7+ (norm %g)",
                                (double)norm));
            }
            idx_out[produced * width + b] = (uint8_t)g;
            if (mag_out) mag_out[produced * width + b] = norm;
        }
        produced++;
        if (i + stride < i) break; /* u64 overflow guard */
    }
    free(inp);
    plan_free(&plan);
    *rows_out = produced;
    QO_LEAVE(QO_OK);
}

int qo_spark_fft_text(qo_samples *s, size_t width, uint64_t stride, int has_min, float min, int has_max, float max,
                      char *out, size_t cap, size_t *len_out)
{
    uint64_t rows = 0;
    int rc = qo_spark_rows(s, width, stride, &rows);
    if (rc != QO_OK) return rc;
    uint8_t *idx = malloc((size_t)(rows ? rows : 1) * width);
    uint64_t got = 0;
    /* fft.rs:19 prints the header before anything can fail */
    size_t o = (size_t)snprintf(out, cap, "sparkfft sample_rate=%llu\n", (unsigned long long)qo_sample_rate(s));
    rc = qo_spark_fft(s, width, stride, has_min, min, has_max, max, 0, rows, idx, NULL, &got);
    for (uint64_t r = 0; r < got; r++) {
        if (o + 3 * width + 8 > cap) {
            free(idx);
            return qo_err(QO_E_INVALID_ARG, "text buffer too small");
        }
        o += qo_format_row(idx + r * width, width, out + o);
        out[o++] = '\n';
    }
    free(idx);
    *len_out = o;
    return rc;
}

/* freq_levels, fft.rs:77-101 */
int qo_freq_levels(qo_samples *s, size_t width, uint64_t stride, size_t levels, uint64_t first, uint64_t max_n,
                   uint8_t *vals, uint64_t *total_out)
{
    QO_ENTER();
    if (levels != 2) QO_LEAVE(qo_err(QO_E_LEVELS, "only supporting two levels for now"));
    fft_plan plan;
    plan_init(&plan, width);
    if (stride == 0) {
        plan_free(&plan);
        QO_LEAVE(qo_err(QO_E_ZERO_STRIDE, "attempt to divide by zero"));
    }
    uint64_t total = (s_len(s) - (uint64_t)width) / stride; /* fft.rs:86 */
    *total_out = total;
    qo_cf32 *inp = xcalloc(width, sizeof(qo_cf32));
    uint64_t produced = 0;
    for (uint64_t reading = first; reading < total && produced < max_n; reading++) {
        memset(inp, 0, width * sizeof(qo_cf32));
        int rc = s_read_exact_at(s, reading * stride, inp, width);
        if (rc != QO_OK) { /* .unwrap() */
            free(inp);
            plan_free(&plan);
            QO_LEAVE(rc);
        }
        plan_process(&plan, inp);
        float firsts = 0.0f, seconds = 0.0f; /* Iterator::sum, sequential */
        for (size_t b = 0; b < width / 2; b++) firsts = firsts + hypotf(inp[b].re, inp[b].im);
        for (size_t b = width / 2; b < width; b++) seconds = seconds + hypotf(inp[b].re, inp[b].im);
        vals[produced++] = firsts < seconds ? 0 : 1;
    }
    free(inp);
    plan_free(&plan);
    QO_LEAVE(QO_OK);
}

/* generate_blackman_harris_window, ffts.rs:110-119 */
void qo_blackman_harris(size_t n, float *out)
{
    const float TAU32 = 6.28318530717958647692f; /* std::f32::consts::TAU */
    for (size_t i = 0; i < n; i++) {
        float x = TAU32 * (float)i / (float)(n - 1);
        float value = 0.35875f - 0.48829f * cosf(x) + 0.14128f * cosf(2.0f * x) - 0.01168f * cosf(3.0f * x);
        out[i] = value;
    }
}

/* take_fft, ffts.rs:18-85 (any width, as FftPlanner allows) */
int qo_take_fft(const qo_samples *s, int has_slice, uint64_t start, uint64_t end, size_t width, int blackman_harris,
                size_t output_len, float *out)
{
    QO_ENTER();
    fft_plan plan;
    plan_init_any(&plan, width); /* FftPlanner: any width */
    uint64_t len = s_len(s);
    uint64_t start_sample = has_slice ? start : 0;
    uint64_t end_sample = has_slice ? end : len - (uint64_t)width;
    if (!(end_sample > start_sample)) {
        plan_free(&plan);
        QO_LEAVE(qo_err(QO_E_SLICE, "Invalid slice: end (%llu) must be greater than start (%llu)",
                        (unsigned long long)end_sample, (unsigned long long)start_sample));
    }
    if (!(end_sample < len)) {
        plan_free(&plan);
        QO_LEAVE(qo_err(QO_E_SLICE, "Slice end (%llu) exceeds sample length (%llu)", (unsigned long long)end_sample,
                        (unsigned long long)len));
    }
    uint64_t visible = end_sample - start_sample;
    if (!(visible > (uint64_t)output_len)) {
        plan_free(&plan);
        QO_LEAVE(qo_err(QO_E_VISIBLE, "Visible samples (%llu) must be greater than output length (%zu)",
                        (unsigned long long)visible, output_len));
    }
    double step = (double)visible / (double)output_len;
    qo_cf32 *cbuf = xcalloc(width, sizeof(qo_cf32));
    float *window = NULL;
    if (blackman_harris) {
        window = xcalloc(width, sizeof(float));
        qo_blackman_harris(width, window);
    }
    for (size_t i = 0; i < output_len; i++) {
        uint64_t sample_index = start_sample + f64_as_u64(round(step * (double)i)); /* ffts.rs:60 */
        int rc = s_read_exact_at(s, sample_index, cbuf, width);
        if (rc != QO_OK) {
            free(cbuf);
            free(window);
            plan_free(&plan);
            QO_LEAVE(rc);
        }
        if (window)
            for (size_t j = 0; j < width; j++) { /* Complex<f32> *= f32 */
                cbuf[j].re = cbuf[j].re * window[j];
                cbuf[j].im = cbuf[j].im * window[j];
            }
        plan_process_any(&plan, cbuf);
        for (size_t b = 0; b < width; b++) {
            size_t src = b < width - width / 2 ? b + width / 2 : b - (width - width / 2);
            out[i * width + b] = hypotf(cbuf[src].re, cbuf[src].im);
        }
    }
    free(cbuf);
    free(window);
    plan_free(&plan);
    QO_LEAVE(QO_OK);
}

/* do_write's pull loop, lib.rs:199-210, into memory */
int qo_write_mem(qo_samples *s, size_t chunk, uint64_t first_chunk, uint64_t max_chunks, qo_cf32 *out, uint64_t cap,
                 uint64_t *n_out)
{
    QO_ENTER();
    *n_out = 0;
    qo_cf32 *buf = xcalloc(chunk, sizeof(qo_cf32));
    uint64_t off = first_chunk * (uint64_t)chunk;
    uint64_t done = 0, chunks = 0;
    while (off < s_len(s) && chunks < max_chunks) {
        memset(buf, 0, chunk * sizeof(qo_cf32));
        size_t read = s_read_at(s, off, buf, chunk);
        if (read == 0) { /* assert_ne!(0, read, ...) lib.rs:203 */
            free(buf);
            *n_out = done;
            QO_LEAVE(qo_err(QO_E_WRITE_SHORT, "short read at offset %llu of %llu", (unsigned long long)off,
                            (unsigned long long)s_len(s)));
        }
        off += read;
        if (done + read > cap) {
            free(buf);
            QO_LEAVE(qo_err(QO_E_INVALID_ARG, "output buffer too small"));
        }
        memcpy(out + done, buf, read * sizeof(qo_cf32));
        done += read;
        *n_out = done;
        chunks++;
    }
    free(buf);
    QO_LEAVE(QO_OK);
}

/* do_write, lib.rs:178-213 */
int qo_write_file(qo_samples *s, const char *prefix, int overwrite, char *name_out, size_t name_cap)
{
    QO_ENTER();
    if (strcmp(prefix, "-") == 0) QO_LEAVE(qo_err(QO_E_UNIMPLEMENTED, "not implemented"));
    char name[4096];
    snprintf(name, sizeof name, "%s.sr%llu.cf32", prefix, (unsigned long long)qo_sample_rate(s));
    if (name_out) snprintf(name_out, name_cap, "%s", name);
    int flags = O_WRONLY | (overwrite ? O_CREAT : (O_CREAT | O_EXCL)); /* create vs create_new; no truncate */
    int fd = open(name, flags, 0666);
    if (fd < 0) QO_LEAVE(qo_err(errno == EEXIST ? QO_E_EXISTS : QO_E_IO, "%s: %s", name, strerror(errno)));
    FILE *f = fdopen(fd, "wb");
    qo_cf32 buf[0x1000];
    volatile uint64_t off = 0;
    jmp_buf jb2;
    jmp_buf *outer = tl_jmp;
    tl_jmp = &jb2;
    if (setjmp(jb2)) { /* a panic inside the loop: BufWriter is flushed on unwind */
        tl_jmp = outer;
        fclose(f);
        QO_LEAVE(tl_code);
    }
    while (off < s_len(s)) {
        memset(buf, 0, sizeof buf);
        size_t read = s_read_at(s, off, buf, 0x1000);
        if (read == 0)
            qo_panic(QO_E_WRITE_SHORT, "short read at offset %llu of %llu", (unsigned long long)off,
                     (unsigned long long)s_len(s));
        off += read;
        fwrite(buf, sizeof(qo_cf32), read, f); /* LE f32 re, im */
    }
    tl_jmp = outer;
    fclose(f);
    QO_LEAVE(QO_OK);
}

/* Exhaustive check of the product's divide-free decode (quadrs_b200/csrc/qd_fast.cu div_exact):
 * q0 = x*c, r = fma(-q0, den, x), q = fma(r, c, q0) with c = fl(1/den) must equal x/den for every
 * input the formats can hold.  Returns the number of mismatches. */
int qo_check_div_trick(void)
{
    const float dens[3] = {127.0f, 255.0f, 65535.0f};
    const int los[3] = {-128, 0, -32768}, his[3] = {127, 255, 32767};
    int bad = 0;
    for (int k = 0; k < 3; k++) {
        const float den = dens[k], c = 1.0f / den;
        for (int b = los[k]; b <= his[k]; b++) {
            const float x = (float)b;
            const float q0 = x * c;
            const float r = __builtin_fmaf(-q0, den, x);
            const float q = __builtin_fmaf(r, c, q0);
            if (q != x / den) bad++;
        }
    }
    return bad;
}

/* ------------------------------------------------------------------ */
/* timed CPU baseline                                                  */
/* ------------------------------------------------------------------ */

static qo_samples *build_job_chain(const qo_job *job)
{
    qo_samples *s = qo_from_mem(job->data, job->n_bytes, job->format, job->sample_rate);
    if (!s) return NULL;
    for (uint32_t i = 0; i < job->n_stages; i++) {
        qo_samples *next = NULL;
        int rc = job->stage_kind[i] == 1
                     ? qo_shift(s, job->stage_freq[i], &next)
                     : qo_lowpass(s, (uint64_t)job->stage_freq[i], job->stage_decimate[i], (size_t)job->stage_size[i],
                                  &next);
        if (rc != QO_OK) {
            qo_free(s);
            return NULL;
        }
        s = next;
    }
    return s;
}

typedef struct {
    const qo_job *job;
    uint64_t first, count;
    uint64_t checksum;
    int rc;
} job_slice;

static void *job_thread(void *arg)
{
    job_slice *sl = arg;
    const qo_job *job = sl->job;
    qo_samples *s = build_job_chain(job);
    if (!s) {
        sl->rc = QO_E_INVALID_ARG;
        return NULL;
    }
    uint64_t sum = 0;
    if (job->sink == 0) {
        qo_cf32 *out = malloc(0x1000 * sizeof(qo_cf32));
        for (uint64_t c = 0; c < sl->count; c++) {
            uint64_t n = 0;
            int rc = qo_write_mem(s, 0x1000, sl->first + c, 1, out, 0x1000, &n);
            if (rc != QO_OK && rc != QO_E_WRITE_SHORT) sl->rc = rc;
            for (uint64_t i = 0; i < n; i++) { /* position-weighted, so that swapped samples do not cancel */
                uint32_t a, b;
                memcpy(&a, &out[i].re, 4);
                memcpy(&b, &out[i].im, 4);
                sum += (a + ((uint64_t)b << 1)) * (i + 1);
            }
        }
        free(out);
    } else {
        size_t W = (size_t)job->width;
        uint64_t batch = 64;
        uint8_t *idx = malloc(batch * W);
        for (uint64_t r = 0; r < sl->count; r += batch) {
            uint64_t want = sl->count - r < batch ? sl->count - r : batch, got = 0;
            int rc = qo_spark_fft(s, W, job->stride, job->has_range, job->min, job->has_range, job->max,
                                  sl->first + r, want, idx, NULL, &got);
            if (rc != QO_OK) sl->rc = rc;
            for (uint64_t i = 0; i < got * W; i++) sum += (uint64_t)idx[i] * (i % W + 1); /* weighted by the bin */
        }
        free(idx);
    }
    sl->checksum = sum;
    qo_free(s);
    return NULL;
}

double qo_timed_run(const qo_job *job, uint64_t first_unit, uint64_t n_units, int n_threads, uint64_t *checksum_out)
{
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_units && n_units > 0) n_threads = (int)n_units;
    pthread_t *th = calloc((size_t)n_threads, sizeof *th);
    job_slice *sl = calloc((size_t)n_threads, sizeof *sl);
    uint64_t per = n_units / (uint64_t)n_threads, rem = n_units % (uint64_t)n_threads, cur = first_unit;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; t++) {
        sl[t].job = job;
        sl[t].first = cur;
        sl[t].count = per + ((uint64_t)t < rem ? 1 : 0);
        cur += sl[t].count;
        pthread_create(&th[t], NULL, job_thread, &sl[t]);
    }
    uint64_t sum = 0;
    int bad = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        sum += sl[t].checksum;
        if (sl[t].rc) bad = sl[t].rc;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    free(sl);
    if (checksum_out) *checksum_out = sum;
    if (bad) return -(double)bad;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
