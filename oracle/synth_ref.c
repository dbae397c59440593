/*
 * oracle/synth_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU twin of the library's device-side synthetic IQ generator
 * (quadrs_b200/csrc/qd_synth.cu, qd_synth_fill).  Not part of the reference:
 * quadrs has no synthetic integer source; this exists so the oracle and the
 * GPU path can be fed bit-identical bytes at any absolute sample index
 * (SURVEY.md section 8d).  Integer-only per sample: a u32 phase accumulator
 * indexing a 4096-entry int16 sine table, plus splitmix64 noise keyed by the
 * absolute sample index.
 */
#include "quadrs_oracle.h"

#include <math.h>
#include <string.h>

static int16_t g_sine[4096];
static int g_sine_ready;

static void sine_init(void)
{
    if (g_sine_ready) return;
    for (int i = 0; i < 4096; i++) g_sine[i] = (int16_t)lround(32767.0 * sin(2.0 * M_PI * (double)i / 4096.0));
    g_sine_ready = 1;
}

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline int32_t component(const qo_synth *p, uint64_t n, int q)
{
    int32_t acc = 0;
    for (uint32_t t = 0; t < p->n_tones; t++) {
        if (p->key_period[t] && !((n / p->key_period[t]) & 1)) continue;
        uint32_t ph = (uint32_t)(n * (uint64_t)p->tone_step[t]);
        if (!q) ph += 0x40000000u; /* I = cos, Q = sin */
        int32_t s = g_sine[ph >> 20];
        acc += (p->tone_amp[t] * s) >> 15; /* arithmetic shift */
    }
    if (p->noise_amp > 0) {
        uint64_t h = splitmix64(p->seed ^ (2 * n + (uint64_t)q));
        uint32_t span = 2u * (uint32_t)p->noise_amp + 1u;
        acc += (int32_t)((uint32_t)(h >> 33) % span) - p->noise_amp;
    }
    return acc;
}

static inline int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : v > hi ? hi : v; }

void qo_synth_fill(const qo_synth *p, int format, uint64_t first, uint64_t n_samples, uint8_t *out)
{
    sine_init();
    for (uint64_t k = 0; k < n_samples; k++) {
        uint64_t n = first + k;
        int32_t vi = component(p, n, 0), vq = component(p, n, 1);
        switch (format) {
        case QO_CS8:
            out[2 * k] = (uint8_t)(int8_t)clampi(vi, -128, 127);
            out[2 * k + 1] = (uint8_t)(int8_t)clampi(vq, -128, 127);
            break;
        case QO_CU8:
            out[2 * k] = (uint8_t)clampi(vi + 128, 0, 255);
            out[2 * k + 1] = (uint8_t)clampi(vq + 128, 0, 255);
            break;
        case QO_CS16: {
            int16_t a = (int16_t)clampi(vi, -32768, 32767), b = (int16_t)clampi(vq, -32768, 32767);
            memcpy(out + 4 * k, &a, 2);
            memcpy(out + 4 * k + 2, &b, 2);
            break;
        }
        case QO_CF32: {
            float a = (float)vi * (1.0f / 32768.0f), b = (float)vq * (1.0f / 32768.0f); /* exact */
            memcpy(out + 8 * k, &a, 4);
            memcpy(out + 8 * k + 4, &b, 4);
            break;
        }
        }
    }
}
