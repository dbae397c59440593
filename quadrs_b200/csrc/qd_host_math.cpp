// qd_host_math.cpp -- host-side constants of the chain: filter taps, windows, twiddles, tables.
//
// The reference computes these once per stage on the CPU too (LowPass::new computes its taps
// eagerly, filter.rs:29-31).  Operation order follows the reference exactly; libm is glibc's, the
// same one Rust's std calls.  Built with -ffp-contract=off.
#include <cmath>
#include <cstring>

#include "qd_internal.h"

namespace qd {

uint64_t pair_bytes(int format)
{ // FileFormat::pair_bytes, lib.rs:217-229
    switch (format) {
    case QD_FMT_CF32: return 8;
    case QD_FMT_CS8:
    case QD_FMT_CU8: return 2;
    case QD_FMT_CS16: return 4;
    default: return 0;
    }
}

double shift_ratio(int64_t frequency, uint64_t sample_rate)
{ // shift.rs:28 with TAU = PI * 2. (lib.rs:23)
    const double tau = M_PI * 2.0;
    return tau * static_cast<double>(frequency) / static_cast<double>(sample_rate);
}

void lowpass_taps(uint64_t frequency, uint64_t sample_rate, size_t size, float *out)
{
    // LowPass::new: cutoff_from_frequency(frequency as f64, sr) as f32  (filter.rs:29-31,126-128)
    const float cutoff = static_cast<float>(static_cast<double>(frequency) / static_cast<double>(sample_rate));
    const float pi = 3.14159274101257324219f; // std::f32::consts::PI
    const float n = static_cast<float>(size);
    for (size_t i = 0; i < size; i++) {
        const float fi = static_cast<float>(i);
        // blackman_window, filter.rs:91-94
        const float w = 0.42f - 0.5f * cosf(2.0f * pi * fi / (n - 1.0f)) + 0.08f * cosf(4.0f * pi * fi / (n - 1.0f));
        // sinc(2.0 * cutoff * (i - (size - 1)/2)), filter.rs:87-89,96-97
        const float x = 2.0f * cutoff * (fi - (n - 1.0f) / 2.0f);
        const float xp = x * pi;
        out[i] = (sinf(xp) / xp) * w;
    }
    float sum = 0.0f; // filter.rs:103-104: sequential sum, then divide
    for (size_t i = 0; i < size; i++) sum = sum + out[i];
    for (size_t i = 0; i < size; i++) out[i] = out[i] / sum;
}

void blackman_harris(size_t n, float *out)
{ // generate_blackman_harris_window, ffts.rs:110-119
    const float tau = 6.28318530717958647692f; // std::f32::consts::TAU
    for (size_t i = 0; i < n; i++) {
        const float x = tau * static_cast<float>(i) / static_cast<float>(n - 1);
        out[i] = 0.35875f - 0.48829f * cosf(x) + 0.14128f * cosf(2.0f * x) - 0.01168f * cosf(3.0f * x);
    }
}

void fft_twiddles(size_t n, float *out)
{
    // Our FFT definition (rustfft 6.4.0 is not in the reference tree): w(N, j) = e^{-2 pi i j/N},
    // angle formed in f64 as (-2 pi / N) * j, cos/sin in f64, then rounded to f32.
    const double constant = -2.0 * M_PI / static_cast<double>(n);
    for (size_t j = 0; j < n; j++) {
        const double angle = constant * static_cast<double>(j);
        out[2 * j] = static_cast<float>(cos(angle));
        out[2 * j + 1] = static_cast<float>(sin(angle));
    }
}

void sincos_table(double *out)
{
    // cos/sin(2 pi i / 256) as double-double, from 80-bit long double evaluation.
    const long double tau = 6.283185307179586476925286766559005768L;
    // The second half turn is the exact negation of the first (e^{i(x + pi)} = -e^{ix}): the device reads only
    // entries 0..127 and flips signs (sincos_f64k), the rest is kept for inspection.
    for (int i = 0; i < 128; i++) {
        const long double a = tau * static_cast<long double>(i) / 256.0L;
        long double c = cosl(a), s = sinl(a);
        if (i == 0) { c = 1.0L; s = 0.0L; }
        if (i == 64) { c = 0.0L; s = 1.0L; }
        const double ch = static_cast<double>(c), sh = static_cast<double>(s);
        out[4 * i + 0] = ch;
        out[4 * i + 1] = static_cast<double>(c - static_cast<long double>(ch));
        out[4 * i + 2] = sh;
        out[4 * i + 3] = static_cast<double>(s - static_cast<long double>(sh));
        for (int j = 0; j < 4; j++) out[4 * (i + 128) + j] = -out[4 * i + j];
    }
}

void sine_table_i16(int16_t *out)
{
    for (int i = 0; i < 4096; i++)
        out[i] = static_cast<int16_t>(lround(32767.0 * sin(2.0 * M_PI * static_cast<double>(i) / 4096.0)));
}

bool is_pow2(uint64_t v) { return v && !(v & (v - 1)); }

uint64_t f64_as_u64(double v)
{ // Rust `as u64`: saturating, NaN -> 0
    if (!(v > 0.0)) return 0;
    if (v >= 18446744073709551616.0) return UINT64_MAX;
    return static_cast<uint64_t>(v);
}

} // namespace qd

namespace qd {

// fft.rs:53-60 on the host, with the panic zone reported as 9
static int glyph_host(float norm, float mn, float mx)
{
    if (norm < mn) return 0;
    if (norm >= mx) return 8;
    const float distinction = (mx - mn) / 7.0f;
    const float q = (norm - mn) / distinction;
    if (!(q > 0.0f)) return 1;
    if (q >= 7.0f) return 9;
    return 1 + static_cast<int>(q);
}

// The glyph is a non-decreasing step function of the magnitude in the order ' ', 7 bars, panic zone, full
// block; the magnitude (float)sqrt(s) is a non-decreasing function of s = fl64(re^2 + im^2).  So each
// boundary is ONE f64 threshold on s, found by bisection over bit patterns with the exact host arithmetic.
bool spark_thresholds(float mn, float mx, double thr[9])
{
    auto rank = [&](float norm) {
        const int g = glyph_host(norm, mn, mx);
        return g <= 7 ? g : (g == 9 ? 8 : 9);
    };
    auto f32_from_bits = [](uint32_t b) {
        float f;
        memcpy(&f, &b, 4);
        return f;
    };
    auto f64_from_bits = [](uint64_t b) {
        double d;
        memcpy(&d, &b, 8);
        return d;
    };
    const uint32_t inf32 = 0x7f800000u;
    const uint64_t inf64 = 0x7ff0000000000000ull;
    for (int r = 1; r <= 9; r++) {
        // smallest non-negative float (by bit pattern) with rank >= r, or none
        if (rank(f32_from_bits(inf32)) < r) {
            thr[r - 1] = f64_from_bits(0x7ff8000000000000ull); // NaN: no s compares >= it
            continue;
        }
        uint32_t lo = 0, hi = inf32; // hi satisfies
        if (rank(f32_from_bits(0)) >= r) hi = 0;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (rank(f32_from_bits(mid)) >= r) hi = mid;
            else lo = mid + 1;
        }
        const float nb = f32_from_bits(hi);
        // smallest non-negative double s with (float)sqrt(s) >= nb
        uint64_t a = 0, b = inf64;
        auto ok = [&](uint64_t bits) { return static_cast<float>(sqrt(f64_from_bits(bits))) >= nb; };
        if (ok(0)) b = 0;
        while (a < b) {
            const uint64_t mid = a + (b - a) / 2;
            if (ok(mid)) b = mid;
            else a = mid + 1;
        }
        thr[r - 1] = f64_from_bits(b);
    }
    return true;
}

} // namespace qd
