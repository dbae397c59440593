// qd_synth.cu -- synthetic IQ generator (bench / test input; not part of the reference).
//
// Integer-only per sample and keyed by the absolute sample index, so a CPU twin
// (oracle/synth_ref.c) regenerates any slice bit-identically and each GPU shard can fill its own
// range in place.  u32 phase accumulator -> 4096-entry int16 sine table -> int32 sum,
// plus splitmix64 noise, then clamp and store in the capture format.
#include <algorithm>

#include "qd_internal.h"

namespace qd {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ int synth_component(const qd_synth &p, const int16_t *__restrict__ sine, uint64_t n, int q)
{
    int acc = 0;
    for (uint32_t t = 0; t < p.n_tones; t++) {
        if (p.key_period[t] && !((n / p.key_period[t]) & 1)) continue;
        uint32_t ph = static_cast<uint32_t>(n * static_cast<uint64_t>(p.tone_step[t]));
        if (!q) ph += 0x40000000u; // I = cos, Q = sin
        acc += (p.tone_amp[t] * static_cast<int>(sine[ph >> 20])) >> 15;
    }
    if (p.noise_amp > 0) {
        const uint64_t h = splitmix64(p.seed ^ (2 * n + static_cast<uint64_t>(q)));
        const uint32_t span = 2u * static_cast<uint32_t>(p.noise_amp) + 1u;
        acc += static_cast<int>(static_cast<uint32_t>(h >> 33) % span) - p.noise_amp;
    }
    return acc;
}

__global__ void k_synth_fill(qd_synth p, int format, uint64_t first, uint64_t n_samples,
                             const int16_t *__restrict__ sine_g, void *__restrict__ out)
{
    __shared__ int16_t sine[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sine[i] = sine_g[i];
    __syncthreads();
    const uint64_t step = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n_samples; k += step) {
        const uint64_t n = first + k;
        const int vi = synth_component(p, sine, n, 0), vq = synth_component(p, sine, n, 1);
        switch (format) {
        case QD_FMT_CS8:
            reinterpret_cast<char2 *>(out)[k] =
                make_char2(static_cast<signed char>(max(-128, min(127, vi))), static_cast<signed char>(max(-128, min(127, vq))));
            break;
        case QD_FMT_CU8:
            reinterpret_cast<uchar2 *>(out)[k] = make_uchar2(static_cast<unsigned char>(max(0, min(255, vi + 128))),
                                                             static_cast<unsigned char>(max(0, min(255, vq + 128))));
            break;
        case QD_FMT_CS16:
            reinterpret_cast<short2 *>(out)[k] =
                make_short2(static_cast<short>(max(-32768, min(32767, vi))), static_cast<short>(max(-32768, min(32767, vq))));
            break;
        default:
            reinterpret_cast<float2 *>(out)[k] =
                make_float2(static_cast<float>(vi) * (1.0f / 32768.0f), static_cast<float>(vq) * (1.0f / 32768.0f));
            break;
        }
    }
}

int synth_fill(const qd_synth *p, int format, uint64_t first, uint64_t n, void *d_out, int device, cudaStream_t st)
{
    if (!p || !d_out) return set_error(QD_E_INVALID_ARG, "qd_synth_fill: null argument");
    if (p->n_tones > 8 || !pair_bytes(format)) return set_error(QD_E_INVALID_ARG, "qd_synth_fill: bad tones/format");
    if (n == 0) return QD_OK;
    DeviceCtx *ctx = nullptr;
    QD_TRY(device_ctx(device, &ctx));
    QD_CUDA(cudaSetDevice(device));
    const uint64_t want = (n + 255) / 256;
    const unsigned blocks = static_cast<unsigned>(std::min<uint64_t>(want, static_cast<uint64_t>(ctx->sm_count) * 16));
    k_synth_fill<<<blocks, 256, 0, st>>>(*p, format, first, n, ctx->d_sine_i16, d_out);
    QD_LAUNCHED();
    return QD_OK;
}

} // namespace qd
