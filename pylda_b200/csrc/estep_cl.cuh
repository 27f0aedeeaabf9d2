// Per-document VB E-step kernel for LONG documents: one document per thread-block CLUSTER (sm_100a).
//
// Documents whose n_d x K tile does not fit one CTA's 227 KB of shared memory (28 % of all
// (document, term) pairs at the headline config) used to re-stream their B rows from L2 on every
// fixed-point trip.  Here a cluster of C = 2, 4 or 8 CTAs owns the document: CTA r keeps rows
// [r*npc, (r+1)*npc) resident in ITS shared memory (bulk-async staged once, as in estep_v2),
// computes the column partial sums of its slice, and the K-vector is all-reduced across the
// cluster through distributed shared memory: every owner thread stores its partial into the
// exchange slot of every peer (st.shared::cluster), one barrier.cluster per trip, then every CTA
// sums the C partials in the same order -- so gamma, e and the convergence decision are computed
// redundantly but bit-identically in all CTAs and no second exchange is needed.  The exchange
// slots are double buffered (trip parity), which is what makes one cluster barrier per trip enough.
//
// One CTA = 8 warps = one "group" of estep_v2 with W = 8; static round-robin of documents over
// clusters (documents are sorted by length, so neighbouring clusters get similar work).
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

template <int LK, int J>
__global__ void __launch_bounds__(256) estep_cl(const EParams p) {
    constexpr int W = 8;
    constexpr int LN = 32 / LK;
    constexpr int KPAD = 2 * LK * J;
    constexpr int GT = 256;
    constexpr int U = (KPAD + GT - 1) / GT;
    constexpr int RR = (2 * J <= 16) ? 2 : 1;
    constexpr int MAXC = 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int gt = threadIdx.x;
    const int gw = gt >> 5;
    const int lane = gt & 31;
    const int kl = lane % LK;
    const int nl = lane / LK;
    const int K = p.K, KP = p.KP, ST = p.ST;
    const int KP2 = KP >> 1;
    const uint32_t crank = cluster_ctarank();
    const uint32_t C = cluster_nctarank();

    unsigned char* gs = smem_raw;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(gs);
    double* es = reinterpret_cast<double*>(gs + 16);
    double* spart = reinterpret_cast<double*>(gs + p.off_spart);     // [W][KPAD]
    double* red = reinterpret_cast<double*>(gs + p.off_red);
    double* xbuf = reinterpret_cast<double*>(gs + p.off_gam);        // [2][MAXC][KPAD] exchange slots
    double* cnt = reinterpret_cast<double*>(gs + p.off_cnt);
    double* mwr = reinterpret_cast<double*>(gs + p.off_mwr);
    int* rid = reinterpret_cast<int*>(gs + p.off_rid);
    double* tile = reinterpret_cast<double*>(gs + p.off_tile);

    for (int i = 16 + gt * 8; i < p.group_bytes; i += GT * 8) *reinterpret_cast<double*>(gs + i) = 0.0;
    if (gt == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_arrive();   // every CTA of the cluster has initialised its shared memory
    cluster_wait();

    double alr[U], gamr[U], er[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int k = gt + GT * u;
        alr[u] = (k < K) ? p.alpha[k] : 1.0;
        gamr[u] = 1.0;
        er[u] = 0.0;
    }
    const bool warp_owns = gw * 32 < K;
    uint32_t parity = 0;

    for (int idx = (int)cluster_id_x(); idx < p.ndocs; idx += (int)num_clusters_x()) {
        bulk_wait_read0();
        __syncthreads();     // previous document retired by every thread of this CTA
        cluster_arrive();    // ... and by every CTA of the cluster (its exchange slots are free again);
                             // the matching wait sits just before the first trip, behind the staging
        const int d = p.order[idx];
        const long long base = p.row_ptr[d];
        const int n = (int)(p.row_ptr[d + 1] - base);
        // this CTA's slice of rows
        int npc = (n + (int)C - 1) / (int)C;
        npc = (npc + LN - 1) / LN * LN;
        const int rlo = min(n, (int)crank * npc);
        const int rhi = min(n, rlo + npc);
        const int nloc = rhi - rlo;
        const int npad = (nloc + LN - 1) / LN * LN;
        const int NG = npad / LN;

        if (gt == 0) mbar_expect_tx(mbar, (uint32_t)npad * (uint32_t)KP * 8u);
        for (int r = gt; r < npad; r += GT) {
            const bool real = r < nloc;
            const int id = p.ids[base + rlo + (real ? r : 0)];
            const int c = real ? p.cts[base + rlo + r] : 0;
            rid[r] = id;
            cnt[r] = (double)c;
            mwr[r] = p.mw[id];
            bulk_g2s(tile + (size_t)r * ST, p.Bt + (size_t)id * KP, (uint32_t)KP * 8u, mbar);
        }
        // N_d over the WHOLE document (every CTA sums all counts: n <= a few thousand ints)
        int csum = 0;
        for (int r = gt; r < n; r += GT) csum += p.cts[base + r];
        csum = __reduce_add_sync(0xffffffffu, csum);
        if (lane == 0) red[gw] = (double)csum;
        __syncthreads();
        double Nd = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) Nd += red[w];
        const double g0 = Nd / (double)K;                 // gamma0 = alpha + N_d / K   (:165)
#pragma unroll
        for (int u = 0; u < U; ++u) gamr[u] = alr[u] + g0;
        if (warp_owns) {
#pragma unroll
            for (int u = 0; u < U; ++u) er[u] = exp_digamma(gamr[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                if (k < K) es[k] = er[u];
            }
        }
        mbar_wait(mbar, parity);
        parity ^= 1u;
        __syncthreads();
        cluster_wait();

        double e[2 * J];
        int it = 0;
        const double tolK = p.tol * (double)K;
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
            {
                const double* rowp = tile + (size_t)(gw * LN + nl) * ST + 2 * kl;
                const double* cntp = cnt + gw * LN + nl;
                const size_t gstride = (size_t)W * LN * ST;
                int q = gw;
                for (; q + W * (RR - 1) < NG; q += W * RR, rowp += RR * gstride, cntp += RR * W * LN)
                    rows_accum<LK, J, RR>(rowp, gstride, cntp, W * LN, e, s);
                if (RR > 1) {
                    for (; q < NG; q += W, rowp += gstride, cntp += W * LN)
                        rows_accum<LK, J, 1>(rowp, gstride, cntp, W * LN, e, s);
                }
            }
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
            __syncthreads();
            // CTA partial of the column sums -> exchange slot (trip parity) of every CTA in the cluster
            double* slot = xbuf + (size_t)(it & 1) * MAXC * KPAD;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                if (k < K) {
                    double ss0 = 0.0, ss1 = 0.0;
#pragma unroll
                    for (int q = 0; q < W; q += 2) {
                        ss0 += spart[q * KPAD + k];
                        ss1 += spart[(q + 1) * KPAD + k];
                    }
                    const double mine = ss0 + ss1;
                    for (uint32_t r = 0; r < C; ++r) st_cluster_f64(slot + crank * KPAD + k, r, mine);
                }
            }
            cluster_arrive();
            cluster_wait();
            // every CTA: same C partials, same order -> bit-identical gamma / e / convergence decision
            double gn[U], en[U];
            double dsum = 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                double ss = 0.0;
                if (k < K) {
                    for (uint32_t r = 0; r < C; ++r) ss += slot[r * KPAD + k];
                }
                gn[u] = fma(er[u], ss, alr[u]);                           // :185
                if (k < K) dsum += fabs(gn[u] - gamr[u]);                 // :187
            }
            if (warp_owns) {
#pragma unroll
                for (int u = 0; u < U; ++u) en[u] = exp_digamma(gn[u]);
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) en[u] = 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) gamr[u] = gn[u];                  // :188
            ++it;
            dsum = warp_sum(dsum);
            if (lane == 0) red[gw] = dsum;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                if (k < K) es[k] = en[u];
            }
            __syncthreads();
            dsum = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) dsum += red[w];
            if (dsum <= tolK || it >= p.max_iter) break;                  // :189-190 / :174
#pragma unroll
            for (int u = 0; u < U; ++u) er[u] = en[u];
        }

        // ---- final pass over this CTA's slice: phi from the LAST e ------------------------------
        double lacc = 0.0;
        {
            double* rowp = tile + (size_t)(gw * LN + nl) * ST + 2 * kl;
            for (int r0 = gw * LN; r0 < nloc; r0 += W * LN, rowp += (size_t)W * LN * ST) {
                const int r = r0 + nl;
                const bool ok = r < nloc;
                double b[2 * J];
                const double part = row_dot<LK, J>(rowp, e, b, ok ? min(J, (KP2 - kl + LK - 1) / LK) : 0);
                const double c = cnt[r];
                const double w = ok ? c * rcp_nr(part) : 0.0;
                if (ok && kl == 0) lacc = fma(c, mwr[r] + log(part), lacc);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (ok && kl + LK * j < KP2)
                        *reinterpret_cast<double2*>(rowp + 2 * LK * j) =
                            make_double2(w * b[2 * j] * e[2 * j], w * b[2 * j + 1] * e[2 * j + 1]);   // :207
                }
            }
        }
        fence_async_smem();
        double t1 = lacc, sg = 0.0;
        if (crank == 0) {     // the document-level ELBO pieces and gamma are written once
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                if (k < K) {
                    const double gk = gamr[u];
                    const double ek = er[u];
                    const double dk = gk - alr[u];
                    t1 += lgamma(gk);                                        // :197
                    if (ek > 0.0 && dk != 0.0) t1 -= log(ek) * dk;
                    sg += gk;
                    p.gamma[(size_t)d * K + k] = gk;                         // :212 / :216
                }
            }
        }
        t1 = warp_sum(t1);
        sg = warp_sum(sg);
        if (lane == 0) {
            red[W + 2 * gw] = t1;
            red[W + 2 * gw + 1] = sg;
        }
        __syncthreads();
        for (int r = gt; r < nloc; r += GT)
            bulk_red_add_f64(p.phi_ss + (size_t)rid[r] * KP, tile + (size_t)r * ST, (uint32_t)KP * 8u);
        bulk_commit();
        if (gt == 0) {
            t1 = 0.0;
            sg = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                t1 += red[W + 2 * w];
                sg += red[W + 2 * w + 1];
            }
            // docterm[d] was zeroed by the host; every CTA adds its share
            atomicAdd(p.docterm + d, crank == 0 ? t1 - lgamma(sg) : t1);      // - lgamma(sum_k gamma_k), :197
            if (crank == 0) p.iters[d] = it;
        }
    }
    bulk_wait_read0();
    cluster_arrive();   // nobody leaves while a peer may still address its shared memory
    cluster_wait();
}

}  // namespace pylda
