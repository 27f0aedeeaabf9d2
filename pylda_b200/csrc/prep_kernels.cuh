// Once-per-E-step kernels around the per-document kernel: the E_log_eta producer
// (inferencer.py:15-18, called at variational_bayes.py:152), the B-table build, the ELBO
// reductions, layout transposes and the device M-step (variational_bayes.py:218-226).
// All of them are plain HBM-streaming kernels over K x V doubles.
#pragma once
#include <cuda_runtime.h>
#include "special.cuh"

namespace pylda {

__device__ __forceinline__ double block_sum(double v, double* sh /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (wid == 0) r = warp_sum(r);
    if (threadIdx.x == 0) sh[0] = r;
    __syncthreads();
    r = sh[0];
    return r;
}

// psisum[k] = psi(sum_v eta[k][v])                                  (inferencer.py:18)
// row sums of eta (K, V): grid (K, chunks of V) so that small K still fills the machine; the partial sums go to
// part[k * chunks + c] and k_psi_of_rowsum adds them in a fixed order (bit-reproducible from call to call) and turns
// them into psi(sum_v eta_kv) (inferencer.py:18)
__global__ void k_rowsum(const double* __restrict__ eta, int K, int V, double* __restrict__ part) {
    __shared__ double sh[32];
    const int k = blockIdx.x;
    const int per = (V + gridDim.y - 1) / gridDim.y;
    const int v0 = blockIdx.y * per, v1 = min(V, v0 + per);
    const double* row = eta + (size_t)k * V;
    double a = 0.0;
    for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) a += row[v];
    a = block_sum(a, sh);
    if (threadIdx.x == 0) part[(size_t)k * gridDim.y + blockIdx.y] = a;
}
__global__ void k_psi_of_rowsum(const double* __restrict__ part, int K, int chunks, double* __restrict__ psisum,
                                double* __restrict__ rowsum) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) {
        double a = 0.0;
        for (int c = 0; c < chunks; ++c) a += part[(size_t)k * chunks + c];
        psisum[k] = digamma_pos(a);
        if (rowsum) rowsum[k] = a;
    }
}

__global__ void k_elog_transpose(const double* __restrict__ eta, const double* __restrict__ psisum, int K, int V,
                                 int KP, double* __restrict__ Elt) {
    __shared__ double t[32][33];
    const int v0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, v = v0 + threadIdx.x;
        if (k < K && v < V) t[i][threadIdx.x] = digamma_pos(eta[(size_t)k * V + v]) - psisum[k];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int v = v0 + i, k = k0 + threadIdx.x;
        if (v < V && k < K) Elt[(size_t)v * KP + k] = t[threadIdx.x][i];
    }
}

// One warp per word: m_w = max_k Elt, Bt = exp(Elt - m_w); zero the padding column and the
// word's row of the statistics accumulator (variational_bayes.py:147).  flat_part[block] = sum over the block's
// words of sum_k Bt[w,k] (fixed order: bit-reproducible) -- how flat the model is across topics, see launch_estep.
__global__ void k_build_B(const double* __restrict__ Elt, int K, int V, int KP, double* __restrict__ Bt,
                          double* __restrict__ mw, double* __restrict__ phi, double* __restrict__ flat_part) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double tot = 0.0;
    for (int v = warp; v < V; v += nwarps) {
        const double* row = Elt + (size_t)v * KP;
        double m = -1.0e308;
        for (int k = lane; k < K; k += 32) m = fmax(m, row[k]);
        m = warp_max(m);
        double sb = 0.0;
        for (int k = lane; k < KP; k += 32) {
            const double b = (k < K) ? exp(row[k] - m) : 0.0;
            Bt[(size_t)v * KP + k] = b;
            phi[(size_t)v * KP + k] = 0.0;
            sb += b;
        }
        tot += warp_sum(sb);
        if (lane == 0) mw[v] = m;
    }
    __shared__ double wtot[32];
    if (lane == 0) wtot[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += wtot[w];
        flat_part[blockIdx.x] = a;
    }
}

// Column-wise logsumexp over v of Elt[v][k] (held-out branch, variational_bayes.py:155):
// stage 1 = per-(k, v-chunk) online (max, sum); stage 2 = combine.
__global__ void k_lse_partial(const double* __restrict__ Elt, int K, int V, int KP, double* pm, double* ps) {
    const int k = blockIdx.x * 32 + (threadIdx.x & 31);
    const int wy = threadIdx.x >> 5, nwy = blockDim.x >> 5;
    const int chunk = (V + gridDim.y - 1) / gridDim.y;
    const int va = blockIdx.y * chunk, vb = min(V, va + chunk);
    double m = -1.0e308, s = 0.0;
    if (k < K) {
        for (int v = va + wy; v < vb; v += nwy) {
            const double x = Elt[(size_t)v * KP + k];
            if (x > m) {
                s = s * exp(m - x) + 1.0;
                m = x;
            } else {
                s += exp(x - m);
            }
        }
    }
    __shared__ double sm[8][32], ss[8][32];
    sm[wy][threadIdx.x & 31] = m;
    ss[wy][threadIdx.x & 31] = s;
    __syncthreads();
    if (wy == 0 && k < K) {
        double M = m;
        for (int i = 1; i < nwy; ++i) M = fmax(M, sm[i][threadIdx.x]);
        double S = 0.0;
        for (int i = 0; i < nwy; ++i) S += ss[i][threadIdx.x] * exp(sm[i][threadIdx.x] - M);
        pm[(size_t)blockIdx.y * K + k] = M;
        ps[(size_t)blockIdx.y * K + k] = S;
    }
}
__global__ void k_lse_final(const double* pm, const double* ps, int K, int nchunk, double* lse) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double M = -1.0e308;
    for (int i = 0; i < nchunk; ++i) M = fmax(M, pm[(size_t)i * K + k]);
    double S = 0.0;
    for (int i = 0; i < nchunk; ++i) S += ps[(size_t)i * K + k] * exp(pm[(size_t)i * K + k] - M);
    lse[k] = M + log(S);
}

// Partial sums for the ELBO: per block
//   [0] sum_d docterm   [1] sum_{w,k} phi*Elt   [2] sum_{w,k} phi*(Elt - lse_k)  (held-out, :204)
//   [3] sum_d iters     [4] #docs at the iteration cap     [5] sum_d n_d * iters_d (row-trips: fp64 work)
constexpr int NTERMS = 6;
__global__ void k_reduce_terms(const double* __restrict__ phi, const double* __restrict__ Elt,
                               const double* __restrict__ lse, int K, int V, int KP,
                               const double* __restrict__ docterm, const int* __restrict__ iters,
                               const long long* __restrict__ row_ptr, long long D, int max_iter, int heldout,
                               double* partial) {
    __shared__ double sh[32];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nel = (long long)V * KP;
    for (long long i = t0; i < nel; i += stride) {
        const int k = (int)(i % KP);
        if (k < K) {
            const double ph = phi[i];
            if (ph != 0.0) {
                const double el = Elt[i];
                a1 = fma(ph, el, a1);
                if (heldout) a2 = fma(ph, el - lse[k], a2);
            }
        }
    }
    for (long long i = t0; i < D; i += stride) {
        a0 += docterm[i];
        const int it = iters[i];
        a3 += (double)it;
        a4 += (it >= max_iter) ? 1.0 : 0.0;
        a5 += (double)it * (double)(row_ptr[i + 1] - row_ptr[i]);
    }
    a0 = block_sum(a0, sh);
    a1 = block_sum(a1, sh);
    a2 = block_sum(a2, sh);
    a3 = block_sum(a3, sh);
    a4 = block_sum(a4, sh);
    a5 = block_sum(a5, sh);
    if (threadIdx.x == 0) {
        double* o = partial + (size_t)blockIdx.x * NTERMS;
        o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3; o[4] = a4; o[5] = a5;
    }
}
__global__ void k_reduce_final(const double* partial, int nblocks, int nterms, double* out) {
    __shared__ double sh[32];
    for (int t = 0; t < nterms; ++t) {
        double a = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += blockDim.x) a += partial[(size_t)i * nterms + t];
        a = block_sum(a, sh);
        if (threadIdx.x == 0) out[t] = a;
    }
}

// (V, KP) -> (K, V) for the reference's phi_sufficient_statistics layout (:147)
__global__ void k_transpose_VK_to_KV(const double* __restrict__ in, int K, int V, int KP, double* __restrict__ out) {
    __shared__ double t[32][33];
    const int v0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int v = v0 + i, k = k0 + threadIdx.x;
        if (v < V && k < K) t[i][threadIdx.x] = in[(size_t)v * KP + k];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, v = v0 + threadIdx.x;
        if (k < K && v < V) out[(size_t)k * V + v] = t[threadIdx.x][i];
    }
}

// alpha statistics (variational_bayes.py:232-233): out_partial[block][k] = sum over the block's
// documents of psi(gamma_dk) - psi(sum_k gamma_dk).  One warp per document.
__global__ void k_alpha_ss(const double* __restrict__ gamma, long long D, int K, double* partial) {
    extern __shared__ double acc[];   // [warps][K]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* my = acc + (size_t)wid * K;
    for (int k = lane; k < K; k += 32) my[k] = 0.0;
    __syncwarp();
    for (long long d = (long long)blockIdx.x * nw + wid; d < D; d += (long long)gridDim.x * nw) {
        const double* g = gamma + (size_t)d * K;
        double s = 0.0;
        for (int k = lane; k < K; k += 32) s += g[k];
        s = warp_sum(s);
        const double ps = digamma_pos(s);
        for (int k = lane; k < K; k += 32) my[k] += digamma_pos(g[k]) - ps;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double a = 0.0;
        for (int w = 0; w < nw; ++w) a += acc[(size_t)w * K + k];
        partial[(size_t)blockIdx.x * K + k] = a;
    }
}

// Device M-step (variational_bayes.py:222-226), one block per topic row:
// rowterm[k] = sum_v lgamma(eta_kv) - lgamma(sum_v eta_kv) from the OLD eta, then
// eta_kv <- phi_KV[k][v] + alpha_beta.
// Device M-step (variational_bayes.py:222-226) straight from the (V, KP) statistics: a 32 x 32 tile of phi is read
// along k, turned through shared memory, and meets the (K, V) rows of eta along v -- no (K, V) copy of the
// statistics.  Grid (chunks of V, K / 32): a block walks the tiles of its V-chunk and leaves, per topic,
// part[k * chunks + c] = sum_v lgamma(eta_old) and part[(K + k) * chunks + c] = sum_v eta_old over the chunk;
// k_mstep_final adds the chunks in a fixed order.  eta <- phi + alpha_beta.
__global__ void k_mstep_tiled(double* __restrict__ eta, const double* __restrict__ phi, int K, int V, int KP, double alpha_beta,
                              double* __restrict__ part) {
    __shared__ double tile[32][33];
    const int chunks = gridDim.x;
    const int tiles = (V + 31) / 32, per = (tiles + chunks - 1) / chunks;
    const int k0 = blockIdx.y * 32;
    double lg[4] = {0.0, 0.0, 0.0, 0.0}, sm[4] = {0.0, 0.0, 0.0, 0.0};      // blockDim = (32, 8): 4 topics per thread
    for (int t = blockIdx.x * per; t < min(tiles, (int)(blockIdx.x + 1) * per); ++t) {
        const int v0 = t * 32;
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += 8) {                       // rows of the tile = words
            const int v = v0 + i, k = k0 + threadIdx.x;
            tile[i][threadIdx.x] = (v < V && k < K) ? phi[(size_t)v * KP + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                     // rows of the result = topics
            const int i = threadIdx.y + 8 * q;
            const int k = k0 + i, v = v0 + threadIdx.x;
            if (k < K && v < V) {
                const size_t o = (size_t)k * V + v;
                const double x = eta[o];
                lg[q] += lgamma(x);
                sm[q] += x;
                eta[o] = tile[threadIdx.x][i] + alpha_beta;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = k0 + threadIdx.y + 8 * q;
        const double a = warp_sum(lg[q]), b = warp_sum(sm[q]);
        if (threadIdx.x == 0 && k < K) {
            part[(size_t)k * chunks + blockIdx.x] = a;
            part[(size_t)(K + k) * chunks + blockIdx.x] = b;
        }
    }
}
__global__ void k_mstep_final(const double* __restrict__ part, int K, int chunks, double* __restrict__ rowterm) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) {
        double a = 0.0, b = 0.0;
        for (int c = 0; c < chunks; ++c) {
            a += part[(size_t)k * chunks + c];
            b += part[(size_t)(K + k) * chunks + c];
        }
        rowterm[k] = a - lgamma(b);
    }
}

__global__ void k_special(int which, long long n, const double* __restrict__ x, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    out[i] = which == 0 ? digamma_pos(v) : which == 1 ? exp_digamma_shifted(v, 0.0) : which == 2 ? lgamma(v)
           : which == 3 ? rcp_nr(v) : exp_digamma(v);
}

// exp(psi(alpha_k)) with the device function the E-step kernels use: the e of a topic eliminated as dead
// (gamma_k == alpha_k), read by the narrow stages (estep_narrow.cuh)
__global__ void k_e_dead(const double* __restrict__ alpha, int K, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) out[k] = exp_digamma(alpha[k]);
}

// phi_ss[w,k] += ed_k B[w,k] wsum[w]: the statistics of the topics eliminated as dead, for all documents that
// finished in the narrow stages (estep_narrow.cuh).  One thread per topic pair, words in the grid's y/loop.
__global__ void k_dead_phi(const double* __restrict__ Bt, const double* __restrict__ wsum, const double* __restrict__ e_dead,
                           int K, int V, int KP, double* __restrict__ phi) {
    const int wpb = blockDim.x / 64;                       // words per block pass (64 threads per word)
    const int t = threadIdx.x & 63, wl = threadIdx.x >> 6;
    for (long long w = (long long)blockIdx.x * wpb + wl; w < V; w += (long long)gridDim.x * wpb) {
        const double ws = wsum[w];
        if (ws == 0.0) continue;
        for (int k = t; k < K; k += 64) {
            const size_t o = (size_t)w * KP + k;
            phi[o] = fma(e_dead[k] * ws, Bt[o], phi[o]);
        }
    }
}

// gamma rows of the documents order[0 .. n) from the device array into the device alias of the caller's page-locked
// buffer (one warp per document, 256-byte stores): the rows that were not final yet when the early copy of gamma
// started (capi.cu, estep_resident_impl)
__global__ void k_copy_rows(const int* __restrict__ order, long long n, int K, const double* __restrict__ src,
                            double* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const size_t o = (size_t)order[i] * K;
        for (int k = lane; k < K; k += 32) dst[o + k] = src[o + k];
    }
}

}  // namespace pylda
