// Instantiations of the narrow-stage E-step kernels (estep_narrow.cuh): NC live columns, G lanes per document.
#include "estep_narrow.cuh"
#include "estep_dispatch.h"
namespace pylda {
const void* estep_narrow_lookup(int NC, int G, int RPL, int MINB) {
#define PYLDA_CASE(NN, GG, RR) \
    if (NC == NN && G == GG && RPL == RR && MINB == 2) return (const void*)estep_narrow<NN, GG, RR, 2>; \
    if (NC == NN && G == GG && RPL == RR && MINB == 3 && RR * NN <= 48) return (const void*)estep_narrow<NN, GG, RR * NN <= 48 ? RR : 48 / NN, 3>;
    PYLDA_CASE(32, 32, 3)
    PYLDA_CASE(16, 8, 3) PYLDA_CASE(16, 16, 3) PYLDA_CASE(16, 32, 3) PYLDA_CASE(16, 32, 6)
    PYLDA_CASE(8, 4, 6) PYLDA_CASE(8, 8, 6) PYLDA_CASE(8, 16, 6) PYLDA_CASE(8, 32, 6)
#undef PYLDA_CASE
    return nullptr;
}
}  // namespace pylda
