// Instantiations of the streaming long-document E-step kernel (estep_v2.cuh, estep_stream), one per lane shape.
#include "estep_v2.cuh"
#include "estep_dispatch.h"
namespace pylda {
// streaming kernel (estep_v2.cuh, estep_stream): one per lane shape
const void* estep_stream_lookup(int LK, int J, int mode) {
#define PYLDA_CASE(LL, JJ) if (LK == LL && J == JJ) return mode == 1 ? (const void*)estep_stream<LL, JJ, 1> : mode == 2 ? (const void*)estep_stream<LL, JJ, 2> : (const void*)estep_stream<LL, JJ, 0>;
#define PYLDA_ROW(LL) PYLDA_CASE(LL, 5) PYLDA_CASE(LL, 7) PYLDA_CASE(LL, 8)
    PYLDA_ROW(1) PYLDA_ROW(2) PYLDA_ROW(4) PYLDA_ROW(8) PYLDA_ROW(16) PYLDA_ROW(32)
    PYLDA_CASE(4, 13)
    PYLDA_CASE(32, 16)
#undef PYLDA_ROW
#undef PYLDA_CASE
    return nullptr;
}
}  // namespace pylda
