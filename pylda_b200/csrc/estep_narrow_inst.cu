// Instantiations of the narrow-stage E-step kernels (estep_narrow.cuh): NC live columns, G lanes per document.
#include "estep_narrow.cuh"
#include "estep_dispatch.h"
namespace pylda {
const void* estep_narrow_lookup(int NC, int G) {
#define PYLDA_CASE(NN, GG) if (NC == NN && G == GG) return (const void*)estep_narrow<NN, GG>;
    PYLDA_CASE(16, 8) PYLDA_CASE(16, 16) PYLDA_CASE(16, 32)
    PYLDA_CASE(8, 4) PYLDA_CASE(8, 8) PYLDA_CASE(8, 16) PYLDA_CASE(8, 32)
#undef PYLDA_CASE
    return nullptr;
}
}  // namespace pylda
