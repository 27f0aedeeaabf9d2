// Instantiation of the compact stage for long documents (estep_longc.cuh).
#include "estep_longc.cuh"
#include "estep_dispatch.h"
namespace pylda {
const void* estep_longc_lookup(int NC) {
    if (NC == 32) return (const void*)estep_longc<32>;
    return nullptr;
}
}  // namespace pylda
