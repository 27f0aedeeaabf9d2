// Common definitions of the per-document VB E-step kernels (sm_100a) -- the hot path of
// /root/reference/variational_bayes.py:159-207, restated in "product form":
//
//   B[w][k]  = exp(E_log_eta[k][w] - m_w),  m_w = max_k E_log_eta[k][w]      (built once per E-step)
//   e_k      = exp(psi(gamma_k))                                              (:177, psi term)
//   norm_n   = sum_k B[w_n][k] e_k           (= exp(logsumexp_k(log_phi) - m_w), :182)
//   gamma_k' = alpha_k + e_k * sum_n (c_n / norm_n) B[w_n][k]                 (:185)
//   stop when mean_k |gamma' - gamma| <= tol or after max_iter trips          (:187-190)
//   phi_ss[w_n][k] += c_n B[w_n][k] e_k / norm_n      with the LAST e (phi lags gamma by one step, :207)
//   entropy/ELBO pieces per document (see DESIGN.md "ELBO regrouping")        (:195-199)
//
// This header holds the kernel parameter block and the PTX helpers (mbarrier, bulk-async copy / reduce-add);
// the kernels live in estep_rt.cuh (register tile), estep_v2.cuh (shared-memory tile, streaming),
// estep_narrow.cuh (few live topics), estep_hy.cuh (cluster, opt-in) and estep_sweep.cuh (precision sweep).
// (The first-generation kernel that used to live here was retired in round 2: the streaming kernel now walks
// documents of any length in chunks.)
//
// Lane layout inside a warp: LK lanes along topics x LN = 32/LK lanes along term rows.
// Lane kl owns topic PAIRS p = kl + LK*j, j < J (16-byte aligned -> LDS.128/LDG.128).
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
    int park_long;                   // documents of 193 .. park_long terms are handed to estep_longc at <= 32 live topics (0: off)
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

}  // namespace pylda
