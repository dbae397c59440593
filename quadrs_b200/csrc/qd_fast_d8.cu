// qd_fast_d8.cu -- fk_fir instantiated for decimate 8: 128 threads x 4 outputs (see qd_fir_kernel.cuh)
#include "qd_fir_kernel.cuh"

namespace qd {

int launch_fir_d8(Chain &c, const FirArgs &a, const FirTaps &t, bool exact) { return launch_fir_dr<8, 4, 128>(c, a, t, exact); }

} // namespace qd
