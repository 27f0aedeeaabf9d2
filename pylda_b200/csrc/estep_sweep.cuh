// Precision-sweep kernel (BASELINE.json configs[3]: "nips.88-05, K=200, fp32 vs fp64 gamma/phi tolerance sweep").
//
// The same fixed point as every other generation (reference variational_bayes.py:174-207 in product form) with
// the STORAGE type of the B tile (TS) and the ARITHMETIC type of the trips (TA) as template parameters:
//   <double, double>  control (must agree with the product kernels to rounding)
//   <float,  double>  tile stored in fp32, every operation in fp64
//   <float,  float>   pure fp32: tile, e, norms, weights, column sums and gamma in fp32 (exp(psi) is evaluated in
//                     fp64 and rounded; the ELBO pieces are accumulated in fp64 from the fp32 quantities)
// It is a measurement tool, not a product path (PYLDA_PRECISION selects it; default is the fp64 product path):
// one warp per document, rows streamed from the (V, KP) table, a lane owns topics lane, lane + 32, ... (K <= 256).
// BASELINE.md section 2 predicts the outcome on the CPU (fp32 tile: gamma 1.7e-5 at K = 100, pure fp32: 5e-5 --
// both fail the 1e-5 bar); scripts/config4_sweep.py reproduces the table on the device.
#pragma once
#include "estep_kernel.cuh"

namespace pylda {

template <typename T>
__device__ __forceinline__ T warp_sum_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename TS, typename TA>
__global__ void __launch_bounds__(256) estep_sweep(const EParams p, const TS* __restrict__ Bs) {
    constexpr int UM = 8;
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int K = p.K, KP = p.KP;
    const TA tolK = (TA)(p.tol * (double)K);
    for (int idx = wg; idx < p.ndocs; idx += nw) {
        const int d = p.order[idx];
        const long long base = p.row_ptr[d];
        const int n = (int)(p.row_ptr[d + 1] - base);
        int csum = 0;
        for (int r = lane; r < n; r += 32) csum += p.cts[base + r];
        csum = __reduce_add_sync(0xffffffffu, csum);
        TA al[UM], g[UM], e[UM];
#pragma unroll
        for (int u = 0; u < UM; ++u) {
            const int k = lane + 32 * u;
            al[u] = (k < K) ? (TA)p.alpha[k] : (TA)1;
            g[u] = al[u] + (TA)((double)csum / (double)K);              // :165
            e[u] = (k < K) ? (TA)exp_digamma((double)g[u]) : (TA)0;
        }
        int it = 0;
        while (true) {
            TA s[UM];
#pragma unroll
            for (int u = 0; u < UM; ++u) s[u] = (TA)0;
            for (int r = 0; r < n; ++r) {
                const int id = p.ids[base + r];
                const TA c = (TA)p.cts[base + r];
                TA b[UM], part = (TA)0;
#pragma unroll
                for (int u = 0; u < UM; ++u) {
                    const int k = lane + 32 * u;
                    b[u] = (k < K) ? (TA)Bs[(size_t)id * KP + k] : (TA)0;
                    part += b[u] * e[u];
                }
                part = warp_sum_t<TA>(part);
                const TA w = c / part;
#pragma unroll
                for (int u = 0; u < UM; ++u) s[u] += w * b[u];
            }
            TA dsum = (TA)0, gn[UM];
#pragma unroll
            for (int u = 0; u < UM; ++u) {
                const int k = lane + 32 * u;
                gn[u] = al[u] + e[u] * s[u];                              // :185
                if (k < K) dsum += (gn[u] > g[u]) ? gn[u] - g[u] : g[u] - gn[u];
                g[u] = gn[u];                                             // :188
            }
            dsum = warp_sum_t<TA>(dsum);
            ++it;
            if (dsum <= tolK || it >= p.max_iter) break;                  // :189-190 / :174
#pragma unroll
            for (int u = 0; u < UM; ++u) {
                const int k = lane + 32 * u;
                e[u] = (k < K) ? (TA)exp_digamma((double)g[u]) : (TA)0;
            }
        }
        // phi from the LAST e (phi lags gamma by one trip), ELBO pieces in fp64 from the TA quantities
        double lacc = 0.0;
        for (int r = 0; r < n; ++r) {
            const int id = p.ids[base + r];
            const TA c = (TA)p.cts[base + r];
            TA b[UM], part = (TA)0;
#pragma unroll
            for (int u = 0; u < UM; ++u) {
                const int k = lane + 32 * u;
                b[u] = (k < K) ? (TA)Bs[(size_t)id * KP + k] : (TA)0;
                part += b[u] * e[u];
            }
            part = warp_sum_t<TA>(part);
            const TA w = c / part;
            if (lane == 0) lacc += (double)c * (p.mw[id] + log((double)part));
#pragma unroll
            for (int u = 0; u < UM; ++u) {
                const int k = lane + 32 * u;
                if (k < K) atomicAdd(p.phi_ss + (size_t)id * KP + k, (double)(w * b[u] * e[u]));   // :207
            }
        }
        double t1 = lacc, sg = 0.0;
#pragma unroll
        for (int u = 0; u < UM; ++u) {
            const int k = lane + 32 * u;
            if (k < K) {
                const double gk = (double)g[u], ek = (double)e[u], dk = gk - (double)al[u];
                t1 += lgamma(gk);                                          // :197
                if (ek > 0.0 && dk != 0.0) t1 -= log(ek) * dk;
                sg += gk;
                p.gamma[(size_t)d * K + k] = gk;                           // :212
            }
        }
        t1 = warp_sum(t1);
        sg = warp_sum(sg);
        if (lane == 0) {
            p.docterm[d] = t1 - lgamma(sg);
            p.iters[d] = it;
        }
    }
}

template <typename TS>
__global__ void k_convert_table(const double* __restrict__ src, size_t n, TS* __restrict__ dst) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (TS)src[i];
}

}  // namespace pylda
