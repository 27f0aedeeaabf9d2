// Instantiations of the compact stage for long documents (estep_longc.cuh).
#include "estep_longc.cuh"
#include "estep_dispatch.h"
namespace pylda {
const void* estep_longc_lookup(int NC, int ctas_per_sm) {
    if (NC == 32 && ctas_per_sm == 2) return (const void*)estep_longc<32, 4, 2>;
    if (NC == 32 && ctas_per_sm == 3) return (const void*)estep_longc<32, 2, 3>;
    return nullptr;
}
}  // namespace pylda
