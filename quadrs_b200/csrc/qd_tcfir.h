// qd_tcfir.h -- the tensor-core FIR (fk_tcfir, qd_tcfir.cu) as the planner sees it.
#pragma once

#include <cstdint>
#include <vector>

namespace qd {

struct Chain;

// Row geometry of a filter of L taps at decimation D (rows of 64 samples at absolute positions)
struct TcGeom {
    uint32_t OPR;  // outputs whose first sample lies in a row: 64 / D
    uint32_t NOUT; // outputs a row's samples take part in
    uint32_t NH;   // columns of one tap half: (re, im) of NOUT partial outputs, padded
    uint32_t N;    // UMMA N = 2 * NH (tap halves hi | lo)
    uint32_t DMAX; // rows that meet in one output
    uint32_t XP;   // pitch of the partials exchange in shared memory (floats, odd)
    int32_t c0;    // column i' of row b is output OPR*b + c0 + i'
    int32_t cown;  // own output t of row b is output OPR*b + cown + t
};
bool tcfir_geometry(uint32_t L, uint32_t D, TcGeom *g);
void tcfir_b_image(const TcGeom &g, const float *taps, uint32_t L, uint32_t D, double ratio_sum, std::vector<uint8_t> &img, float *s_hi, float *s_lo);
// untruncated outputs [g0, g1) of `shift* | lowpass` over a cs8 capture, FAST arithmetic, into d_out[g - g0]
int launch_tcfir(Chain &c, const TcGeom &g, const uint8_t *d_bimg, float s_hi, float s_lo, uint32_t L, uint32_t D, int n_shift,
                 const double *ratios, const uint8_t *d_src, uint64_t src_base, uint64_t src_end, uint64_t g0, uint64_t g1, float2 *d_out);

} // namespace qd
