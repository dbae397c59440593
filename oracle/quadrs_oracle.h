/*
 * oracle/quadrs_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the quadrs streaming IQ DSP chain (decode -> shift ->
 * lowpass/decimate -> sparkfft / bucket / take_fft / write, plus gen), written
 * from the reference's behaviour with every function citing the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or the timed CPU baseline -- never on the product path.
 *
 * PARITY STATUS: the reference is Rust and no Rust toolchain exists in the
 * build image, so this restatement cannot be diffed against a real quadrs
 * binary.  It is pinned by (1) the README OOK known-answer walkthrough
 * (README.md:113-187) on examples/cupboard-superdec.sr400.cf32, (2) the
 * config-1 row structure of README.md:90-94, (3) closed forms derived from
 * src/filter.rs.  The FFT arithmetic of rustfft 6.4.0 (Cargo.lock:3208-3219)
 * is absent from /root/reference: FFT PARITY IS UNPINNED beyond what (1)/(2)
 * constrain; the FFT here is our own radix-4 DIT definition, bounded against a
 * complex128 DFT.
 */
#ifndef QUADRS_ORACLE_H
#define QUADRS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } qo_cf32;

/* FileFormat, src/lib.rs:61-74 */
enum { QO_CF32 = 0, QO_CS8 = 1, QO_CU8 = 2, QO_CS16 = 3 };

/* Status codes: shared numbering with include/quadrs_gpu.h (QD_E_*), so parity
 * tests can compare error behaviour one to one. */
enum {
    QO_OK = 0,
    QO_E_INVALID_ARG = 1,
    QO_E_SHIFT_NYQUIST = 2,  /* shift.rs:20-23 */
    QO_E_ZERO_RATE = 3,      /* shift.rs:24, gen.rs:19 */
    QO_E_OFFSET_EOF = 4,     /* samples.rs:74 */
    QO_E_SHORT_INPUT = 5,    /* filter.rs:46,76 */
    QO_E_SHORT_READ = 6,     /* samples.rs:20-25 */
    QO_E_FFT_WIDTH = 7,      /* rustfft Radix4::new power-of-two assert */
    QO_E_GLYPH_RANGE = 8,    /* fft.rs:59 graph[7] out of bounds */
    QO_E_LEVELS = 9,         /* fft.rs:83 */
    QO_E_SLICE = 10,         /* ffts.rs:32-40 */
    QO_E_VISIBLE = 11,       /* ffts.rs:44-48 */
    QO_E_GEN_ARGS = 12,      /* gen.rs:18-20 */
    QO_E_WRITE_SHORT = 13,   /* lib.rs:203 */
    QO_E_IO = 14,
    QO_E_UNIMPLEMENTED = 17, /* lib.rs:180 */
    QO_E_EXISTS = 18,        /* lib.rs:186-192 create_new */
    QO_E_NOMEM = 19,
    QO_E_ZERO_STRIDE = 20    /* fft.rs:65 would loop forever */
};

typedef struct qo_samples qo_samples;

const char *qo_last_error(void);

/* ---- graph construction (Operation::exec arms, src/lib.rs:89-121) ---- */
qo_samples *qo_from_mem(const uint8_t *data, uint64_t n_bytes, int format, uint64_t sample_rate);
qo_samples *qo_from_mem_window(const uint8_t *data, uint64_t n_bytes, int format, uint64_t sample_rate,
                               uint64_t base_sample, uint64_t total_samples); /* test helper */
qo_samples *qo_from_file(const char *path, int format, uint64_t sample_rate);
int qo_gen(const int64_t *cos_hz, size_t n_cos, uint64_t sample_rate, double seconds, qo_samples **out);
int qo_shift(qo_samples *inner, int64_t frequency, qo_samples **out);
int qo_lowpass(qo_samples *inner, uint64_t frequency, uint64_t decimate, size_t size, qo_samples **out);
void qo_free(qo_samples *s);

/* ---- trait Samples, src/samples.rs:11-28 ---- */
int qo_len(const qo_samples *s, uint64_t *out);
uint64_t qo_sample_rate(const qo_samples *s);
int qo_read_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n, size_t *got);
int qo_read_exact_at(const qo_samples *s, uint64_t off, qo_cf32 *buf, size_t n);

/* 0 (default): literal complex_convolve (filter.rs:107-124), every undecimated
 * output.  1: evaluate only the kept outputs with the same per-output op order
 * (bit-identical, tested); used to make large parity cases finish in seconds. */
void qo_set_kept_only_convolve(int on);

/* ---- sinks ---- */
/* spark_fft, src/fft.rs:12-69.  Rows [first_row, first_row+max_rows) of the
 * reference's row sequence are produced; idx_out[r*width + b] in 0..8 (0 ' ',
 * 1..7 the seven bars, 8 full block), display (fftshifted) order; mag_out
 * (nullable) = hypotf per bin in the same order.  rows_out = rows produced. */
int qo_spark_fft(qo_samples *s, size_t width, uint64_t stride, int has_min, float min, int has_max, float max,
                 uint64_t first_row, uint64_t max_rows, uint8_t *idx_out, float *mag_out, uint64_t *rows_out);
/* number of rows the reference loop prints (fft.rs:27-28,65) */
int qo_spark_rows(const qo_samples *s, size_t width, uint64_t stride, uint64_t *rows);
/* exact stdout bytes of spark_fft (fft.rs:19,63): header + rows, UTF-8 */
int qo_spark_fft_text(qo_samples *s, size_t width, uint64_t stride, int has_min, float min, int has_max, float max,
                      char *out, size_t cap, size_t *len_out);
/* freq_levels, src/fft.rs:77-101: vals[first..first+max_n), total_out = total */
int qo_freq_levels(qo_samples *s, size_t width, uint64_t stride, size_t levels, uint64_t first, uint64_t max_n,
                   uint8_t *vals, uint64_t *total_out);
/* take_fft, src/ffts.rs:18-85: out[output_len * width] */
int qo_take_fft(const qo_samples *s, int has_slice, uint64_t start, uint64_t end, size_t width, int blackman_harris,
                size_t output_len, float *out);
/* do_write, src/lib.rs:178-213, into memory.  Chunks [first_chunk,
 * first_chunk+max_chunks) of `chunk`(=0x1000) samples.  Returns
 * QO_E_WRITE_SHORT where the reference's assert_ne! at lib.rs:203 fires (data
 * before it is still delivered, as in the reference). */
int qo_write_mem(qo_samples *s, size_t chunk, uint64_t first_chunk, uint64_t max_chunks, qo_cf32 *out, uint64_t cap,
                 uint64_t *n_out);
int qo_write_file(qo_samples *s, const char *prefix, int overwrite, char *name_out, size_t name_cap);

/* ---- pieces exposed for unit tests ---- */
void qo_decode(int format, const uint8_t *in, size_t n_samples, qo_cf32 *out); /* lib.rs:231-255 */
int qo_taps(uint64_t frequency, uint64_t sample_rate, size_t size, float *out); /* filter.rs:29-31,86-105 */
void qo_blackman_harris(size_t n, float *out);                                  /* ffts.rs:110-119 */
int qo_fft(qo_cf32 *buf, size_t n);                 /* our radix-4 DIT definition, forward, unnormalised */
void qo_dft_c128(const qo_cf32 *in, size_t n, double *out_re_im); /* O(n^2) complex128 bound */
double qo_shift_ratio(int64_t frequency, uint64_t sample_rate);   /* shift.rs:28 */
/* glyph index for one magnitude (fft.rs:45,53-60); -1 where the reference panics */
int qo_glyph_index(float norm, float min, float max);
int qo_check_div_trick(void); /* exhaustive check of the product's divide-free decode */
size_t qo_format_row(const uint8_t *idx, size_t width, char *out); /* fft.rs:34-36,63 */

/* ---- synthetic IQ generator (CPU twin of qd_synth_fill; integer-only) ---- */
typedef struct {
    uint64_t seed;
    uint32_t n_tones;
    uint32_t tone_step[8]; /* phase step per sample, units of 2^-32 turns */
    int32_t tone_amp[8];   /* peak amplitude in output LSBs (cf32: units of 2^-15) */
    uint32_t key_period[8]; /* 0 = always on; else tone is on while (n / key_period) is odd */
    int32_t noise_amp;     /* uniform integer noise in [-noise_amp, noise_amp] */
} qo_synth;
void qo_synth_fill(const qo_synth *p, int format, uint64_t first_sample, uint64_t n_samples, uint8_t *out);

/* ---- timed CPU baseline (bench.py cpu_baseline / --impl reference) ----
 * Runs the reference algorithm (lazy pull, literal convolve) for sink units
 * [first_unit, first_unit + n_units) split over n_threads disjoint ranges.
 * sink: 0 = write (chunk 0x1000), 1 = spark_fft(width, stride, min, max).
 * Returns wall seconds, <0 on error. */
typedef struct {
    const uint8_t *data;
    uint64_t n_bytes;
    int format;
    uint64_t sample_rate;
    uint32_t n_stages;
    int32_t stage_kind[8]; /* 1 shift, 2 lowpass */
    int64_t stage_freq[8];
    uint64_t stage_decimate[8];
    uint64_t stage_size[8];
    int sink;
    uint64_t width, stride;
    int has_range;
    float min, max;
} qo_job;
double qo_timed_run(const qo_job *job, uint64_t first_unit, uint64_t n_units, int n_threads, uint64_t *checksum_out);

#ifdef __cplusplus
}
#endif
#endif
