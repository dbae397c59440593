// qd_fast_d16.cu -- fk_fir instantiated for decimate 16: 128 threads x 4 outputs (see qd_fir_kernel.cuh)
#include "qd_fir_kernel.cuh"

namespace qd {

int launch_fir_d16(Chain &c, const FirArgs &a, const FirTaps &t, bool exact) { return launch_fir_dr<16, 4, 128>(c, a, t, exact); }

} // namespace qd
