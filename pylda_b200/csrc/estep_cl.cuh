// Thread-block cluster helpers (sm_100a): cluster rank / size, barrier.cluster, distributed-shared-memory stores.
// Used by the hybrid long-document kernel (estep_hy.cuh).  (The first cluster kernel, which kept the whole tile
// in shared memory, was retired in round 2: estep_hy.cuh keeps its exchange protocol and adds register rows.)
#pragma once
#include "estep_v2.cuh"

namespace pylda {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// store a double into the shared memory of CTA `rank` of this cluster at the same offset as `local`
__device__ __forceinline__ void st_cluster_f64(const void* local, uint32_t rank, double v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}

}  // namespace pylda
