// Instantiations of the hybrid register / shared-memory tile kernel for long documents (estep_hy.cuh), one per
// lane shape; R = register rows per lane (4*J*R registers of tile).
#include "estep_hy.cuh"
#include "estep_dispatch.h"
namespace pylda {
template <int J> struct HyRows { static constexpr int R = (J <= 5) ? 5 : (J <= 7) ? 4 : (J <= 8) ? 3 : 2; };
const void* estep_hy_lookup(int LK, int J, int* rows_per_lane) {
#define PYLDA_CASE(LL, JJ) if (LK == LL && J == JJ) { *rows_per_lane = HyRows<JJ>::R; return (const void*)estep_hy<LL, JJ, HyRows<JJ>::R>; }
#define PYLDA_ROW(LL) PYLDA_CASE(LL, 5) PYLDA_CASE(LL, 7) PYLDA_CASE(LL, 8)
    PYLDA_ROW(1) PYLDA_ROW(2) PYLDA_ROW(4) PYLDA_ROW(8) PYLDA_ROW(16) PYLDA_ROW(32)
    PYLDA_CASE(4, 13)
    PYLDA_CASE(32, 16)
#undef PYLDA_ROW
#undef PYLDA_CASE
    return nullptr;
}
}  // namespace pylda
