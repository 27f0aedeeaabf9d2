// Instantiations of the second-generation per-document E-step kernel for LK = 2 topic-lanes per row.
// Only shapes with at most 4 topics per owner thread (2*LK*J <= 128*W) exist.
#include "estep_v2.cuh"
#include "estep_dispatch.h"
namespace pylda {
const void* estep_v2_lk2(int J, int W, int V) {
    constexpr int LK = 2;
#define PYLDA_CASE_W(JJ, WW) \
    if constexpr (2 * LK * JJ <= 128 * WW) {                                                        \
        if (J == JJ && W == WW && V == 0) return (const void*)estep_v2<LK, JJ, WW, 0>;              \
    }
#define PYLDA_CASE(JJ) PYLDA_CASE_W(JJ, 8)      // the one class in use (8 warps per document); others were tuning variants
    PYLDA_CASE(5)
    PYLDA_CASE(7)
    PYLDA_CASE(8)

#undef PYLDA_CASE
#undef PYLDA_CASE_W
    return nullptr;
}
}  // namespace pylda
