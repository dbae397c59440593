// qd_fast_d32.cu -- fk_fir instantiated for decimate 32: 128 threads x 2 outputs (see qd_fir_kernel.cuh)
#include "qd_fir_kernel.cuh"

namespace qd {

int launch_fir_d32(Chain &c, const FirArgs &a, const FirTaps &t, bool exact) { return launch_fir_dr<32, 2, 128>(c, a, t, exact); }

} // namespace qd
