/*
 * quadrs_gpu.h -- C ABI of libquadrs_gpu.so, the B200 (sm_100a) drop-in for the
 * CPU stages of FauxFaux/quadrs's streaming IQ DSP chain.
 *
 * The reference has no FFI today: its seam is `trait Samples`
 * (src/samples.rs:11-28) and the `Operation::exec` fold (src/lib.rs:83-175).
 * Each entry point below names the reference interface it replaces.  A Rust
 * binding (`extern "C"` block + `impl Samples for GpuChain`) is shown in
 * INTEGRATION.md and shipped, unbuilt, under rust/quadrs-gpu-sys/.
 *
 * Conventions
 *  - plain pointers and sizes only; no CUDA or torch types in any signature
 *    (a CUDA stream is passed as void*, a device pointer as const void*).
 *  - every function returns a qd_status (0 = QD_OK) unless noted; the message
 *    for the calling thread's last failure is qd_last_error().
 *  - conditions under which the reference panics (assert!/expect/index) return
 *    a distinct code instead of unwinding across the ABI, so a Rust wrapper can
 *    re-panic with the reference's message.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with QD_E_CUDA.
 *  - handles may be used from any thread; calls on one chain are serialised by
 *    an internal mutex (the reference's Samples is Sync + Send, samples.rs:11).
 */
#ifndef QUADRS_GPU_H
#define QUADRS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QD_ABI_VERSION 2

typedef struct { float re, im; } qd_cf32; /* num_complex::Complex<f32>, interleaved */

/* FileFormat, src/lib.rs:61-74 */
typedef enum { QD_FMT_CF32 = 0, QD_FMT_CS8 = 1, QD_FMT_CU8 = 2, QD_FMT_CS16 = 3 } qd_format;

typedef enum {
    QD_OK = 0,
    QD_E_INVALID_ARG = 1,
    QD_E_SHIFT_NYQUIST = 2,  /* shift.rs:20-23  "frequency must be under half the sample rate" */
    QD_E_ZERO_RATE = 3,      /* shift.rs:24 */
    QD_E_OFFSET_EOF = 4,     /* samples.rs:74   assert!(off < self.len()) */
    QD_E_SHORT_INPUT = 5,    /* filter.rs:46,76 inner shorter than the filter */
    QD_E_SHORT_READ = 6,     /* samples.rs:20-25 read_exact_at: wanted != got (an Err, not a panic) */
    QD_E_FFT_WIDTH = 7,      /* fft.rs:25 Radix4::new needs a power of two */
    QD_E_GLYPH_RANGE = 8,    /* fft.rs:59 graph[7]: index out of bounds */
    QD_E_LEVELS = 9,         /* fft.rs:83 "only supporting two levels for now" */
    QD_E_SLICE = 10,         /* ffts.rs:32-40 */
    QD_E_VISIBLE = 11,       /* ffts.rs:44-48 (an Err) */
    QD_E_GEN_ARGS = 12,      /* gen.rs:18-20 (an Err) */
    QD_E_WRITE_SHORT = 13,   /* lib.rs:203 assert_ne!(0, read): fires AFTER all readable data is delivered */
    QD_E_IO = 14,
    QD_E_CUDA = 15,          /* CUDA runtime/driver error, or no device / extension unusable */
    QD_E_NOT_RESIDENT = 16,  /* a shard was asked for samples outside [base_sample, base_sample + n) */
    QD_E_UNIMPLEMENTED = 17, /* lib.rs:180 write to "-" */
    QD_E_EXISTS = 18,        /* lib.rs:186-192 create_new on an existing file */
    QD_E_NOMEM = 19,
    QD_E_ZERO_STRIDE = 20    /* fft.rs:65 would never terminate; fft.rs:86 divides by zero */
} qd_status;

/* ---- sources: Operation::From / Operation::Gen, src/lib.rs:26-29,54-58,89-101 ---- */
typedef enum {
    QD_SRC_HOST_MEM = 0,   /* raw capture bytes in host memory (what SampleFile preads, samples.rs:72-93) */
    QD_SRC_DEVICE_MEM = 1, /* the same bytes already resident in HBM: pointer aligned to one sample, and the
                              allocation readable from the 16-byte boundary at or before the first sample up to
                              the next 16-byte boundary past the last one (tiles are fetched by 16-byte-granular
                              bulk copies and vector loads; any cudaMalloc block, or a sample-aligned view into
                              one, satisfies it) */
    QD_SRC_FILE = 2,       /* path; the library preads it */
    QD_SRC_GEN = 3         /* gen.rs */
} qd_source_kind;

typedef struct {
    int32_t kind;         /* qd_source_kind */
    int32_t format;       /* qd_format (ignored for GEN) */
    uint64_t sample_rate; /* FileDetails.sample_rate / Gen.sample_rate */
    const void *data;     /* HOST_MEM / DEVICE_MEM: first byte of sample `base_sample` */
    uint64_t n_bytes;     /* bytes at `data` (FILE: 0 = whole file) */
    const char *path;     /* FILE */
    /* Sharding: `data` holds samples [base_sample, base_sample + n_bytes/pair_bytes) of a
     * logical capture that is total_samples long.  0/0 = the buffer is the whole capture.  All phase
     * and end-of-file arithmetic uses absolute indices, so a shard reproduces the unsharded result. */
    uint64_t base_sample;
    uint64_t total_samples;
    /* GEN (gen.rs:10-27) */
    double gen_seconds;
    const int64_t *gen_cos;
    uint64_t gen_n_cos;
} qd_source;

/* ---- stages: Operation::Shift / Operation::LowPass, src/lib.rs:31-38 ---- */
typedef enum { QD_STAGE_SHIFT = 1, QD_STAGE_LOWPASS = 2 } qd_stage_kind;
typedef struct {
    int32_t kind;       /* qd_stage_kind */
    int32_t reserved;
    int64_t frequency;  /* Shift.frequency (i64) / LowPass.frequency (u64) */
    uint64_t decimate;  /* LowPass.decimate */
    uint64_t size;      /* LowPass.size = number of taps (args.rs:161-166: 2*power, default 40) */
} qd_stage;

typedef enum { QD_SPACE_HOST = 0, QD_SPACE_DEVICE = 1 } qd_space;

/* Arithmetic mode.  EXACT reproduces the reference's operation order (f64 phase product per sample,
 * mul-then-add FIR, non-contracted FFT): bit-identical to the CPU oracle by construction.  FAST keeps
 * integer decode bit-exact but uses FMA and block-anchored phase: within 1e-5 relative on cs8/cf32
 * inputs; NOT within 1e-5 on the offset formats cu8/cs16, for which it is refused. */
typedef enum { QD_PRECISION_EXACT = 0, QD_PRECISION_FAST = 1 } qd_precision;

typedef struct qd_chain qd_chain;

const char *qd_last_error(void);        /* thread-local, never NULL */
int qd_abi_version(void);
int qd_device_count(int *count);        /* QD_E_CUDA when no usable device */
uint64_t qd_kernel_launches(void);      /* kernels launched by this library so far (process-wide) */
const char *qd_status_name(int status);

/* Builds the lazy graph: From/Gen then the stages in order (the fold of quadrs.rs:48-56).
 * Performs Shift::new / LowPass::new / Gen::new checks (shift.rs:20-24, gen.rs:18-20). */
int qd_chain_create(const qd_source *src, const qd_stage *stages, size_t n_stages, int device, qd_chain **out);
/* The same graph sharded by sample range over several GPUs of one box, driven by ONE host process -- the
 * reference's caller is one process folding commands into one sink (quadrs.rs:48-56 -> fft.rs:27-66 /
 * lib.rs:178-213).  Every sink call on the handle cuts its unit range (sparkfft rows, write chunks, bucket
 * windows, take_fft rows) into n_dev contiguous parts; device i evaluates part i on its own host thread and
 * stream set, staging only the raw samples its units touch (filter-tap / FFT-window halo included), and its
 * results land at their place in the caller's ONE output buffer or file.  Phase, truncation and end-of-capture
 * arithmetic use absolute sample indices, so there is no collective and the result equals the one-device
 * chain's bit for bit (EXACT and FAST).  Sources: HOST_MEM, FILE, GEN (a DEVICE_MEM capture lives on one
 * device: n_dev must be 1).  Outputs must be host buffers.  devices[] may name a device more than once.
 * Worker threads bind themselves to the CPUs local to their GPU (sysfs local_cpulist of its PCI function)
 * before they allocate pinned staging, so that staging lands on the GPU's NUMA node. */
int qd_chain_create_sharded(const qd_source *src, const qd_stage *stages, size_t n_stages, const int *devices,
                            size_t n_dev, qd_chain **out);
/* number of devices a chain runs on (1 for qd_chain_create) */
int qd_chain_n_devices(const qd_chain *c, size_t *n_dev);
void qd_chain_destroy(qd_chain *c);
/* Runs the chain's kernels and copies on the caller's cudaStream_t (NULL is the legacy default
 * stream).  A new chain owns a private non-blocking stream until this is called. */
int qd_chain_set_stream(qd_chain *c, void *cuda_stream);
int qd_chain_set_precision(qd_chain *c, int precision);
int qd_chain_synchronize(qd_chain *c);
/* Tuning knobs (tests and benches): "use_fast" 0/1 (0 forces the general unit-local executor), "fuse_stft" 0/1/2 (sparkfft inside the
 * filter kernel: never / back-to-back windows / also overlapping windows and two-stage chains),
 * "use_tc" 0/1 (FAST precision over cs8 captures: 1, the default, runs the filter on the tensor cores where the chain's shape allows;
 * 0 keeps the CUDA-core kernel), "glyph_lin" 0/1 (sparkfft bucket indices through the proven linear form of the square root
 * where the range allows it: 1, the default; 0 keeps the per-glyph thresholds; results are identical),
 * "fir_carry" 0/1 (long filters: a CTA carries the samples two consecutive tiles share instead of decoding them twice: 1, the
 * default; results are identical), "segment_bytes" raw bytes staged per pipelined segment for host/file sources,
 * "scratch_budget" bytes of device scratch the general executor may use per batch. */
int qd_chain_set_option(qd_chain *c, const char *key, int64_t value);

/* Bench instrumentation: when enabled, the kernels of every sink / read call on this chain are
 * bracketed by CUDA events on the chain's stream.  read: synchronises, then reports how many bracketed
 * regions ran, their summed device time, and the name of the dominant kernel of the last call. */
int qd_chain_profile(qd_chain *c, int enable);
int qd_chain_profile_read(qd_chain *c, uint64_t *regions, double *total_ms, char *kernel_name, size_t cap);

/* Samples::len / sample_rate (samples.rs:12-13), including LowPass's +1 over-report (filter.rs:45-48)
 * and integer sr/decimate (filter.rs:50-52). */
int qd_chain_len(const qd_chain *c, uint64_t *len);
int qd_chain_sample_rate(const qd_chain *c, uint64_t *rate);
/* number of taps and the f32 taps of stage i (filter.rs:86-105), for inspection */
int qd_chain_taps(const qd_chain *c, size_t stage, float *out, size_t cap, size_t *n);

/* Samples::read_at (samples.rs:15): result equals the reference's for this exact (off, n), including
 * the zero-truncated filter tail of LowPass::read_at (filter.rs:68-80).  *got = samples produced. */
int qd_chain_read_at(qd_chain *c, uint64_t off, qd_cf32 *buf, size_t n, int space, size_t *got);
/* Samples::read_exact_at (samples.rs:17-27): QD_E_SHORT_READ when got != n */
int qd_chain_read_exact_at(qd_chain *c, uint64_t off, qd_cf32 *buf, size_t n, int space);

/* spark_fft (fft.rs:12-69).  rows: number of rows the reference prints. */
int qd_sparkfft_rows(const qd_chain *c, size_t width, uint64_t stride, uint64_t *rows);
/* Rows [first_row, first_row + n_rows).  idx_out[r*width + b]: 0 ' ', 1..7 = "▁▂▃▄▅▆▇", 8 '█', display
 * (fftshifted) order (fft.rs:48-60).  mag_out (nullable): hypot per bin, same order.  has_range=0 uses
 * the reference defaults 0.08 / 1.0 (fft.rs:22-23). */
int qd_sparkfft(qd_chain *c, size_t width, uint64_t stride, int has_range, float min, float max, uint64_t first_row,
                uint64_t n_rows, uint8_t *idx_out, float *mag_out, int space, uint64_t *rows_out);
/* Text helpers: one row "│...│" (fft.rs:63) without newline; returns bytes written. */
size_t qd_format_row(const uint8_t *idx, size_t width, char *out, size_t cap);

/* freq_levels (fft.rs:77-101): vals[first .. first+n) each 0/1; *total = (len - width)/stride */
int qd_freq_levels(qd_chain *c, size_t width, uint64_t stride, size_t levels, uint64_t first, uint64_t n,
                   uint8_t *vals, int space, uint64_t *total);

/* take_fft (ffts.rs:18-85): out[output_len * width] fftshifted magnitudes; windowing 0 Rectangular,
 * 1 BlackmanHarris (ffts.rs:12-16,110-119).  Any width 1..16384, as FftPlanner allows: powers of two run
 * the radix-4 FFT, other widths a direct DFT with f64 accumulation. */
int qd_take_fft(qd_chain *c, int has_slice, uint64_t start, uint64_t end, size_t width, int windowing,
                size_t output_len, float *out, int space);

/* do_write's pull loop (lib.rs:199-210): chunks [first_chunk, first_chunk + n_chunks) of `chunk`
 * samples (the reference uses 0x1000) land contiguously in out[]; *n_out = samples written.  Returns
 * QD_E_WRITE_SHORT where the reference's assert_ne! fires -- the data before it is still in out[]. */
int qd_write_cf32(qd_chain *c, size_t chunk, uint64_t first_chunk, uint64_t n_chunks, qd_cf32 *out, uint64_t cap,
                  int space, uint64_t *n_out);
/* do_write (lib.rs:178-213): "{prefix}.sr{rate}.cf32", create_new unless overwrite */
int qd_write_file(qd_chain *c, const char *prefix, int overwrite, char *name_out, size_t name_cap);

/* ---- sharding by sample range: pure host arithmetic, no device needed ---- */
typedef struct {
    uint64_t first_unit, n_units;     /* sink units (sparkfft rows / write chunks) owned by this shard */
    uint64_t first_sample, n_samples; /* raw source samples the shard must hold (halo included) */
} qd_shard;
/* sink_kind: 0 write(chunk = unit_len), 1 sparkfft(width = unit_len, stride), 2 freq_levels */
int qd_shard_plan(const qd_source *src, const qd_stage *stages, size_t n_stages, int sink_kind, uint64_t unit_len,
                  uint64_t stride, uint32_t n_shards, uint32_t shard, qd_shard *out);

/* ---- synthetic IQ (bench/test input; integer-only, keyed by absolute sample index) ---- */
typedef struct {
    uint64_t seed;
    uint32_t n_tones;
    uint32_t tone_step[8];  /* phase step per sample in 2^-32 turns */
    int32_t tone_amp[8];    /* peak amplitude in output LSBs (cf32: units of 2^-15) */
    uint32_t key_period[8]; /* 0 = always on; else on while (n / key_period) is odd */
    int32_t noise_amp;
} qd_synth;
int qd_synth_fill(const qd_synth *p, int format, uint64_t first_sample, uint64_t n_samples, void *device_out,
                  int device, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* QUADRS_GPU_H */
