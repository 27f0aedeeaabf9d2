// Instantiations of the register-tile per-document E-step kernel for LK = 8 topic-lanes per row.
// Only shapes with at most 4 topics per owner thread (2*LK*J <= 128*W) exist.  R = 0 / NWARPS = 0 select
// the default variant (RtRows<J>::R rows per lane, 8 warps per CTA); *rows_per_lane returns its R.
#include "estep_rt.cuh"
#include "estep_dispatch.h"
namespace pylda {
const void* estep_rt_lk8(int J, int W, int R, int NWARPS, int* rows_per_lane) {
    constexpr int LK = 8;
#define PYLDA_CASE_W(JJ, WW) \
    if constexpr (2 * LK * JJ <= 128 * WW) {                                                                   \
        if (J == JJ && W == WW && R == 0) {                                                                    \
            *rows_per_lane = RtRows<JJ>::R;                                                                    \
            return (const void*)estep_rt<LK, JJ, WW, RtRows<JJ>::R, 8>;                                        \
        }                                                                                                      \
    }
#define PYLDA_CASE(JJ) PYLDA_CASE_W(JJ, 1) PYLDA_CASE_W(JJ, 2) PYLDA_CASE_W(JJ, 4) PYLDA_CASE_W(JJ, 8)
    PYLDA_CASE(5)
    PYLDA_CASE(7)
    PYLDA_CASE(8)
#undef PYLDA_CASE
#undef PYLDA_CASE_W
    *rows_per_lane = R;
    return nullptr;
}
}  // namespace pylda
