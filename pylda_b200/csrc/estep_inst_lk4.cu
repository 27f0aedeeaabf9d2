// Instantiations of the per-document E-step kernel for LK = 4 topic-lanes per row.
#include "estep_kernel.cuh"
#include "estep_dispatch.h"
namespace pylda {
const void* estep_kernel_lk4(int J, bool resident) {
#define PYLDA_CASE(JJ) if (J == JJ) return resident ? (const void*)estep_kernel<4, JJ, true> : (const void*)estep_kernel<4, JJ, false>;
    PYLDA_CASE(5)
    PYLDA_CASE(7)
    PYLDA_CASE(8)

#undef PYLDA_CASE
    return nullptr;
}
}  // namespace pylda
