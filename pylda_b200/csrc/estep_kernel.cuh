// Per-document VB E-step kernel (sm_100a) -- the hot path of
// /root/reference/variational_bayes.py:159-207, restated in "product form":
//
//   B[w][k]  = exp(E_log_eta[k][w] - m_w),  m_w = max_k E_log_eta[k][w]      (built once per E-step)
//   e_k      = exp(psi(gamma_k) - c)                                          (:177, psi term)
//   norm_n   = sum_k B[w_n][k] e_k           (= exp(logsumexp_k(log_phi) - m_w - c), :182)
//   gamma_k' = alpha_k + e_k * sum_n (c_n / norm_n) B[w_n][k]                 (:185)
//   stop when mean_k |gamma' - gamma| <= tol or after max_iter trips          (:187-190)
//   phi_ss[w_n][k] += c_n B[w_n][k] e_k / norm_n      with the LAST e (phi lags gamma by one step, :207)
//   entropy/ELBO pieces per document (see DESIGN.md "ELBO regrouping")        (:195-199)
//
// Work decomposition: a "group" of W warps (W = 1,2,4,8) owns one document at a time and
// pulls the next one from a per-class atomic queue (documents are pre-sorted by length,
// classes are chosen so that the document's n_d x K tile of B fits the group's share of
// shared memory).  The tile is staged ONCE per document by bulk-async copies
// (cp.async.bulk -> UBLKCP, one per term row, completion on an mbarrier) and re-used for
// every fixed-point trip; phi rows are formed in place and accumulated into the global
// statistics by bulk reduce-add (cp.reduce.async.bulk ... add.f64 -> UBLKRED), one per row.
// Documents whose tile cannot fit use the streaming instantiation (RES = false) that
// re-reads B rows from L2/HBM every trip and scatters with red.global.add.f64.
//
// Lane layout inside a warp: LK lanes along topics x LN = 32/LK lanes along term rows.
// Lane kl owns topic PAIRS p = kl + LK*j, j < J (16-byte aligned -> LDS.128/LDG.128).
// Row sums (over k) need log2(LK) shuffle steps; column sums (over n) are accumulated in
// registers across all rows the lane visits and reduced once per trip.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "special.cuh"

namespace pylda {

struct EParams {
    // corpus (CSR) -- variational_bayes.py:98-130 packed
    const long long* __restrict__ row_ptr;
    const int* __restrict__ ids;
    const int* __restrict__ cts;
    // class work queue
    const int* __restrict__ order;   // document ids of this class, longest first
    int ndocs;
    int* counter;                    // atomic queue head
    // model tables (V x KP, KP = K rounded up to even)
    const double* __restrict__ Bt;
    const double* __restrict__ mw;   // (V,) per-word max of E_log_eta
    const double* __restrict__ alpha;
    double alpha_max;
    // outputs
    double* gamma;                   // (D, K)
    double* phi_ss;                  // (V, KP)
    double* docterm;                 // (D,)
    int* iters;                      // (D,)
    int K, KP, ST;                   // topics, table row stride, smem tile row stride (doubles)
    int max_iter;
    double tol;
    // group geometry
    int W;                           // warps per group
    int nmax;                        // tile capacity in rows (RES only)
    int compact;                     // estep_rt: dead-topic elimination enabled
    int* revived;                    // estep_rt: counter of documents in which an eliminated topic came back
    // hand-over to the narrow stages (estep_narrow.cuh); park_nc = 0: off, 16 / 8: live-topic threshold
    int park_nc;
    int* park_rec;                   // PARK_REC ints per document
    double* park_gam;                // PARK_GAM doubles per document
    int* park_lists;                 // PARK_LISTS lists of park_cap documents
    int* park_counts;
    int park_cap;
    int group_bytes;                 // bytes of shared memory per group
    int off_groups;                  // byte offset of group 0 (after the CTA-wide alpha copy)
    int off_gam, off_spart, off_red, off_cnt, off_mwr, off_rid, off_tile;   // within a group
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk async copy (TMA engine, non-tensor form); bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk async reduce-add of f64 elements
__device__ __forceinline__ void bulk_red_add_f64(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void group_bar(int W, int g) {
    if (W == 1) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(W * 32) : "memory");
    }
}

template <int U>
__device__ __forceinline__ void update_e(double* es, const double* gam, int K, int gt, int GT, double negc) {
    for (int k0 = gt; k0 < K; k0 += U * GT) {
        double x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u * GT;
            x[u] = (k < K) ? gam[k] : 1.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = exp_digamma_shifted(x[u], negc);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u * GT;
            if (k < K) es[k] = x[u];
        }
    }
}

template <int LK, int J, bool RES>
__global__ void __launch_bounds__(256) estep_kernel(const EParams p) {
    constexpr int LN = 32 / LK;
    constexpr int KPAD = 2 * LK * J;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int W = p.W;
    const int GT = W * 32;
    const int tid = threadIdx.x;
    const int g = tid / GT;
    const int gt = tid - g * GT;
    const int gw = gt >> 5;
    const int lane = tid & 31;
    const int kl = lane % LK;
    const int nl = lane / LK;
    const int K = p.K, KP = p.KP, ST = p.ST;
    const int KP2 = KP >> 1;

    double* alpha_s = reinterpret_cast<double*>(smem_raw);
    unsigned char* gs = smem_raw + p.off_groups + (size_t)g * p.group_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(gs);
    int* cur = reinterpret_cast<int*>(gs + 8);
    double* es = reinterpret_cast<double*>(gs + 16);
    double* gam = reinterpret_cast<double*>(gs + p.off_gam);
    double* spart = reinterpret_cast<double*>(gs + p.off_spart);
    double* red = reinterpret_cast<double*>(gs + p.off_red);
    double* cnt = reinterpret_cast<double*>(gs + p.off_cnt);
    double* mwr = reinterpret_cast<double*>(gs + p.off_mwr);
    int* rid = reinterpret_cast<int*>(gs + p.off_rid);
    double* tile = reinterpret_cast<double*>(gs + p.off_tile);

    for (int k = tid; k < KPAD; k += blockDim.x) alpha_s[k] = (k < K) ? p.alpha[k] : 0.0;
    for (int k = gt; k < KPAD; k += GT) {
        es[k] = 0.0;
        gam[k] = 1.0;
    }
    if (RES && gt == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    const int ucase = (K + GT - 1) / GT;   // topics per owner thread

    while (true) {
        // ---- fetch the next document of this class -------------------------------------
        group_bar(W, g);   // previous document fully retired (shared memory reuse)
        if (gt == 0) *cur = atomicAdd(p.counter, 1);
        group_bar(W, g);
        const int idx = *cur;
        if (idx >= p.ndocs) break;
        const int d = p.order[idx];
        const long long base = p.row_ptr[d];
        const int n = (int)(p.row_ptr[d + 1] - base);

        // ---- stage rows: ids / counts / m_w, and (RES) the B tile by bulk-async copies ----
        int csum = 0;
        if (RES) {
            if (gt == 0) mbar_expect_tx(mbar, (uint32_t)n * (uint32_t)KP * 8u);
            for (int r = gt; r < n; r += GT) {
                const int id = p.ids[base + r];
                const int c = p.cts[base + r];
                rid[r] = id;
                cnt[r] = (double)c;
                mwr[r] = p.mw[id];
                csum += c;
                bulk_g2s(tile + (size_t)r * ST, p.Bt + (size_t)id * KP, (uint32_t)KP * 8u, mbar);
            }
        } else {
            for (int r = gt; r < n; r += GT) csum += p.cts[base + r];
        }
        csum = __reduce_add_sync(0xffffffffu, csum);
        if (lane == 0) red[gw] = (double)csum;
        group_bar(W, g);
        double Nd = 0.0;
        for (int w = 0; w < W; ++w) Nd += red[w];
        // gamma0 = alpha + N_d / K                               (variational_bayes.py:165)
        const double g0 = Nd / (double)K;
        for (int k = gt; k < K; k += GT) gam[k] = alpha_s[k] + g0;
        double negc = -digamma_rough(p.alpha_max + g0);
        group_bar(W, g);
        if (ucase <= 1) update_e<1>(es, gam, K, gt, GT, negc);
        else if (ucase == 2) update_e<2>(es, gam, K, gt, GT, negc);
        else update_e<4>(es, gam, K, gt, GT, negc);
        if (RES) {
            mbar_wait(mbar, parity);
            parity ^= 1u;
        }
        group_bar(W, g);

        // ---- fixed-point trips                                  (variational_bayes.py:174-190)
        double e[2 * J];
        int it = 0;
        while (true) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const double2 v = *reinterpret_cast<const double2*>(es + 2 * (kl + LK * j));
                e[2 * j] = v.x;
                e[2 * j + 1] = v.y;
            }
            double s[2 * J];
#pragma unroll
            for (int i = 0; i < 2 * J; ++i) s[i] = 0.0;

            for (int r0 = gw * LN; r0 < n; r0 += W * LN) {
                const int r = r0 + nl;
                const bool ok = r < n;
                const double* rowp;
                double c = 0.0;
                if (RES) {
                    rowp = tile + (size_t)(ok ? r : 0) * ST;
                    if (ok) c = cnt[r];
                } else {
                    int id = 0;
                    if (ok) {
                        id = p.ids[base + r];
                        c = (double)p.cts[base + r];
                    }
                    rowp = p.Bt + (size_t)id * KP;
                }
                double b[2 * J];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int pi = kl + LK * j;
                    double2 v = make_double2(0.0, 0.0);
                    if (ok && pi < KP2) v = *reinterpret_cast<const double2*>(rowp + 2 * pi);
                    b[2 * j] = v.x;
                    b[2 * j + 1] = v.y;
                }
                double pa = 0.0, pb = 0.0;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    pa = fma(b[2 * j], e[2 * j], pa);
                    pb = fma(b[2 * j + 1], e[2 * j + 1], pb);
                }
                double part = pa + pb;
#pragma unroll
                for (int o = 1; o < LK; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                const double w = ok ? c / part : 0.0;
#pragma unroll
                for (int i = 0; i < 2 * J; ++i) s[i] = fma(w, b[i], s[i]);
            }
            // column sums: reduce over the LN row-lanes, then over the group's warps
#pragma unroll
            for (int o = LK; o < 32; o <<= 1) {
#pragma unroll
                for (int i = 0; i < 2 * J; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
            }
            if (nl == 0) {
#pragma unroll
                for (int j = 0; j < J; ++j)
                    *reinterpret_cast<double2*>(spart + gw * KPAD + 2 * (kl + LK * j)) =
                        make_double2(s[2 * j], s[2 * j + 1]);
            }
            group_bar(W, g);
            double dsum = 0.0, gmx = 0.0;
            for (int k = gt; k < K; k += GT) {
                double ss = 0.0;
                for (int w = 0; w < W; ++w) ss += spart[w * KPAD + k];
                const double gn = fma(es[k], ss, alpha_s[k]);      // :185
                dsum += fabs(gn - gam[k]);                          // :187
                gam[k] = gn;                                        // :188
                gmx = fmax(gmx, gn);
            }
            dsum = warp_sum(dsum);
            gmx = warp_max(gmx);
            if (lane == 0) {
                red[2 * gw] = dsum;
                red[2 * gw + 1] = gmx;
            }
            group_bar(W, g);
            double dt = 0.0, gm = 0.0;
            for (int w = 0; w < W; ++w) {
                dt += red[2 * w];
                gm = fmax(gm, red[2 * w + 1]);
            }
            ++it;
            if (dt / (double)K <= p.tol || it >= p.max_iter) break;   // :189-190 / :174
            negc = -digamma_rough(gm);
            if (ucase <= 1) update_e<1>(es, gam, K, gt, GT, negc);
            else if (ucase == 2) update_e<2>(es, gam, K, gt, GT, negc);
            else update_e<4>(es, gam, K, gt, GT, negc);
            group_bar(W, g);
        }

        // ---- final pass: phi from the LAST e (es unchanged since the last trip) ------------
        double lacc = 0.0;
        for (int r0 = gw * LN; r0 < n; r0 += W * LN) {
            const int r = r0 + nl;
            const bool ok = r < n;
            double* rowp_s = nullptr;
            const double* rowp;
            double c = 0.0, mwv = 0.0;
            int id = 0;
            if (RES) {
                rowp_s = tile + (size_t)(ok ? r : 0) * ST;
                rowp = rowp_s;
                if (ok) {
                    c = cnt[r];
                    mwv = mwr[r];
                }
            } else {
                if (ok) {
                    id = p.ids[base + r];
                    c = (double)p.cts[base + r];
                    mwv = p.mw[id];
                }
                rowp = p.Bt + (size_t)id * KP;
            }
            double b[2 * J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int pi = kl + LK * j;
                double2 v = make_double2(0.0, 0.0);
                if (ok && pi < KP2) v = *reinterpret_cast<const double2*>(rowp + 2 * pi);
                b[2 * j] = v.x;
                b[2 * j + 1] = v.y;
            }
            double pa = 0.0, pb = 0.0;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                pa = fma(b[2 * j], e[2 * j], pa);
                pb = fma(b[2 * j + 1], e[2 * j + 1], pb);
            }
            double part = pa + pb;
#pragma unroll
            for (int o = 1; o < LK; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            const double w = ok ? c / part : 0.0;
            if (ok && kl == 0) lacc = fma(c, mwv + log(part), lacc);   // sum_n c_n * logsumexp_n
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int pi = kl + LK * j;
                if (ok && pi < KP2) {
                    const double f0 = w * b[2 * j] * e[2 * j];          // c_n phi_nk        (:207)
                    const double f1 = w * b[2 * j + 1] * e[2 * j + 1];
                    if (RES) {
                        *reinterpret_cast<double2*>(rowp_s + 2 * pi) = make_double2(f0, f1);
                    } else {
                        double* dst = p.phi_ss + (size_t)id * KP + 2 * pi;
                        atomicAdd(dst, f0);
                        if (2 * pi + 1 < K) atomicAdd(dst + 1, f1);
                    }
                }
            }
        }
        if (RES) {
            fence_async_smem();   // generic-proxy writes of phi -> visible to the bulk-async engine
        }
        // ---- per-document ELBO pieces and gamma write-back -----------------------------------
        double t1 = lacc, sg = 0.0;
        for (int k = gt; k < K; k += GT) {
            const double gk = gam[k];
            const double ek = es[k];
            const double dk = gk - alpha_s[k];
            t1 += lgamma(gk);                                            // :197
            if (ek > 0.0 && dk != 0.0) t1 -= log(ek) * dk;               // - sum_k psi_k sum_n c_n phi_nk (c cancels)
            sg += gk;
            p.gamma[(size_t)d * K + k] = gk;                             // :212 / :216
        }
        t1 = warp_sum(t1);
        sg = warp_sum(sg);
        group_bar(W, g);   // all phi rows written; red[] free again
        if (lane == 0) {
            red[2 * gw] = t1;
            red[2 * gw + 1] = sg;
        }
        if (RES) {
            for (int r = gt; r < n; r += GT)
                bulk_red_add_f64(p.phi_ss + (size_t)rid[r] * KP, tile + (size_t)r * ST, (uint32_t)KP * 8u);
            bulk_commit();
        }
        group_bar(W, g);
        if (gt == 0) {
            double a = 0.0, b2 = 0.0;
            for (int w = 0; w < W; ++w) {
                a += red[2 * w];
                b2 += red[2 * w + 1];
            }
            p.docterm[d] = a - lgamma(b2);                               // - lgamma(sum_k gamma_k), :197
            p.iters[d] = it;
        }
        if (RES) bulk_wait_read0();   // tile may be overwritten after the next group barrier
    }
}

}  // namespace pylda
