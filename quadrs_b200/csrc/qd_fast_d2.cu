// qd_fast_d2.cu -- fk_fir instantiated for decimate 2: 128 threads x 8 outputs (see qd_fir_kernel.cuh)
#include "qd_fir_kernel.cuh"

namespace qd {

int launch_fir_d2(Chain &c, const FirArgs &a, const FirTaps &t, bool exact) { return launch_fir_dr<2, 8, 128>(c, a, t, exact); }

} // namespace qd
