// Per-document VB E-step kernel for LONG documents, hybrid register / shared-memory tile (sm_100a).
//
// Same mathematics as the other generations (reference variational_bayes.py:159-207 in product form).
// Documents above ~190 terms used to re-read their B rows every trip: from shared memory through a
// latency-bound loop (estep_v2, 24 cycles per row-trip and SM) or from L2 (estep_stream: 14.7 TB/s, which IS
// the L2 + L1 limit -- 35 % of the whole E-step for 2 % of the documents).  Long documents keep all topics
// alive for the whole 50 trips, so what they need is on-chip residency of the WHOLE tile:
//   * a CTA of 8 warps keeps R rows x 2J columns per lane in REGISTERS (160 rows at K = 100, loaded straight
//     from the (V, KP) table with LDG.128) plus up to ~210 rows in SHARED memory (bulk-async staged, read once
//     per trip into a two-row-group register window) -- about 370 rows, 300 KB of tile, per SM;
//   * a thread-block CLUSTER of C = 1, 2, 4 or 8 such CTAs owns one document (rows dealt in contiguous
//     slices); the K-vector of column sums is all-reduced through distributed shared memory exactly as in
//     estep_cl.cuh: every owner thread stores its CTA partial into the exchange slot of every peer
//     (st.shared::cluster), ONE barrier.cluster per trip, slots double buffered on trip parity, then every
//     CTA sums the C partials in the same order, so gamma, e and the stop decision are bit-identical
//     everywhere and no second exchange is needed.
// Per trip a lane issues 4 R J DFMA on registers, 4 J per shared-memory row group, J STS.128 for the column
// partials; owners sum W*LN partials.  phi: shared-memory rows in place (as estep_v2), then the register rows
// go through the same region; both leave by bulk reduce-add (UBLKRED), one per row.
#pragma once
#include "estep_cl.cuh"

namespace pylda {

template <int LK, int J, int R>
struct HyCfg {
    static constexpr int W = 8;
    static constexpr int LN = 32 / LK;
    static constexpr int KPAD = 2 * LK * J;
    static constexpr int GT = 256;
    static constexpr int U = (KPAD + GT - 1) / GT;
    static constexpr int NP = W * LN;           // column-partial rows in shared memory
    static constexpr int CAPR = W * LN * R;     // register rows per CTA
    static constexpr int MAXC = 8;
};

// RR row groups of shared-memory rows: norm, weight, column-sum accumulation
template <int LK, int J, int RR>
__device__ __forceinline__ void hy_smem_rows(const double* rowp, size_t gstride, const double* cntp, int cstride,
                                             const double (&e)[2 * J], double (&s)[2 * J]) {
    double b[RR][2 * J];
    double a0[RR], a1[RR];
#pragma unroll
    for (int i = 0; i < RR; ++i) a0[i] = a1[i] = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int i = 0; i < RR; ++i) {
            const double2 v = *reinterpret_cast<const double2*>(rowp + i * gstride + 2 * LK * j);
            b[i][2 * j] = v.x;
            b[i][2 * j + 1] = v.y;
            a0[i] = fma(v.x, e[2 * j], a0[i]);
            a1[i] = fma(v.y, e[2 * j + 1], a1[i]);
        }
    }
    double part[RR];
#pragma unroll
    for (int i = 0; i < RR; ++i) part[i] = a0[i] + a1[i];
#pragma unroll
    for (int o = 1; o < LK; o <<= 1) {
#pragma unroll
        for (int i = 0; i < RR; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], o);
    }
#pragma unroll
    for (int i = 0; i < RR; ++i) {
        const double w = cntp[i * cstride] * rcp_nr(part[i]);      // pad rows: a valid row with count 0
#pragma unroll
        for (int c = 0; c < 2 * J; ++c) s[c] = fma(w, b[i][c], s[c]);
    }
}

template <int LK, int J, int R>
__global__ void __launch_bounds__(256) estep_hy(const EParams p) {
    using C_ = HyCfg<LK, J, R>;
    constexpr int W = C_::W, LN = C_::LN, KPAD = C_::KPAD, GT = C_::GT, U = C_::U, NP = C_::NP, CAPR = C_::CAPR;
    constexpr int RSTEP = W * LN;
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
    double* es2 = reinterpret_cast<double*>(gs + 16);                 // [2][KPAD], trip parity
    double* spart = reinterpret_cast<double*>(gs + p.off_spart);     // [NP][KPAD]
    double* red = reinterpret_cast<double*>(gs + p.off_red);
    double* xbuf = reinterpret_cast<double*>(gs + p.off_gam);        // [2][C][KPAD] exchange slots
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
    if (C > 1) {
        cluster_arrive();   // every CTA of the cluster has initialised its shared memory
        cluster_wait();
    }

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
    const int rbase = gw * LN + nl;
    const double tolK = p.tol * (double)K;

    for (int idx = (int)cluster_id_x(); idx < p.ndocs; idx += (int)num_clusters_x()) {
        bulk_wait_read0();
        __syncthreads();     // previous document retired by every thread of this CTA
        if (C > 1) cluster_arrive();   // ... and by every CTA of the cluster (exchange slots free again); wait below
        const int d = p.order[idx];
        const long long base = p.row_ptr[d];
        const int n = (int)(p.row_ptr[d + 1] - base);
        // this CTA's slice of rows: the first nR go to registers, the other nS to shared memory
        int npc = (n + (int)C - 1) / (int)C;
        npc = (npc + LN - 1) / LN * LN;
        const int rlo = min(n, (int)crank * npc);
        const int nloc = min(n, rlo + npc) - rlo;
        const int nR = min(nloc, CAPR);
        const int nS = nloc - nR;                          // host guarantees nS <= p.nmax
        const int nSpad = (nS + LN - 1) / LN * LN;
        const int NGS = nSpad / LN;

        if (gt == 0) mbar_expect_tx(mbar, (uint32_t)nSpad * (uint32_t)KP * 8u);
        for (int r = gt; r < nR + nSpad; r += GT) {
            const bool real = r < nloc;
            const int src = real ? r : nR;                 // pad rows (shared-memory part only): a valid row, count 0
            const int id = p.ids[base + rlo + src];
            rid[r] = id;
            cnt[r] = real ? (double)p.cts[base + rlo + r] : 0.0;
            mwr[r] = p.mw[id];
            if (r >= nR) bulk_g2s(tile + (size_t)(r - nR) * ST, p.Bt + (size_t)id * KP, (uint32_t)KP * 8u, mbar);
        }
        // N_d over the WHOLE document (every CTA sums all counts)
        int csum = 0;
        for (int r = gt; r < n; r += GT) csum += p.cts[base + r];
        csum = __reduce_add_sync(0xffffffffu, csum);
        if (lane == 0) red[gw] = (double)csum;
        __syncthreads();     // red[], rid[], cnt[] visible
        double Nd = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) Nd += red[w];
        const double g0 = Nd / (double)K;                  // gamma0 = alpha + N_d / K   (:165)
#pragma unroll
        for (int u = 0; u < U; ++u) gamr[u] = alr[u] + g0;
        if (warp_owns) {
#pragma unroll
            for (int u = 0; u < U; ++u) er[u] = exp_digamma(gamr[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                if (k < K) es2[k] = er[u];
            }
        }
        // register rows: straight from the (V, KP) table
        double b[R][2 * J];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = rbase + i * RSTEP;
            const bool ok = r < nR;
            const double* rowp = p.Bt + (size_t)(ok ? rid[r] : 0) * KP + 2 * kl;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                double2 v = make_double2(0.0, 0.0);
                if (ok && kl + LK * j < KP2) v = *reinterpret_cast<const double2*>(rowp + 2 * LK * j);
                b[i][2 * j] = v.x;
                b[i][2 * j + 1] = v.y;
            }
        }
        mbar_wait(mbar, parity);
        parity ^= 1u;
        __syncthreads();
        if (C > 1) cluster_wait();

        double wv[R], part[R];
        int it = 0;
        // The two row phases of a trip use different pipes (shared-memory rows: LDS-bound; register rows: fp64-bound),
        // so half of the warps (one per scheduler) run them in the opposite order: profiles/r2b_estep_hy_blocks.txt
        // showed all eight warps queueing on the shared-memory pipe at once (LDS latency ~700 cycles).
        const bool smem_first = (gw >> 2) & 1;
        while (true) {
            double e[2 * J];
            {
                const double* es = es2 + (it & 1) * KPAD;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const double2 v = *reinterpret_cast<const double2*>(es + 2 * (kl + LK * j));
                    e[2 * j] = v.x;
                    e[2 * j + 1] = v.y;
                }
            }
            double s[2 * J];
#pragma unroll
            for (int i = 0; i < 2 * J; ++i) s[i] = 0.0;
#pragma unroll 1
            for (int phase = 0; phase < 2; ++phase) {
                if ((phase == 0) == smem_first) {
                    // shared-memory rows: row groups gw, gw + W, ... two at a time
                    const double* rowp = tile + (size_t)rbase * ST + 2 * kl;
                    const double* cntp = cnt + nR + rbase;
                    const size_t gstride = (size_t)RSTEP * ST;
                    int q = gw;
                    for (; q + W < NGS; q += 2 * W, rowp += 2 * gstride, cntp += 2 * RSTEP)
                        hy_smem_rows<LK, J, 2>(rowp, gstride, cntp, RSTEP, e, s);
                    if (q < NGS) hy_smem_rows<LK, J, 1>(rowp, gstride, cntp, RSTEP, e, s);
                } else {
                    // register rows
                    double a0[R], a1[R];
#pragma unroll
                    for (int i = 0; i < R; ++i) a0[i] = a1[i] = 0.0;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
#pragma unroll
                        for (int i = 0; i < R; ++i) {
                            a0[i] = fma(b[i][2 * j], e[2 * j], a0[i]);
                            a1[i] = fma(b[i][2 * j + 1], e[2 * j + 1], a1[i]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < R; ++i) part[i] = a0[i] + a1[i];
#pragma unroll
                    for (int o = 1; o < LK; o <<= 1) {
#pragma unroll
                        for (int i = 0; i < R; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], o);
                    }
                    // (rows past nR: b = 0 gives part = 0; the count is 0 there and the norm is replaced before the
                    // reciprocal, so that the reciprocals stay branch-free and interleave)
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const bool ok = rbase + i * RSTEP < nR;
                        const double c = ok ? cnt[rbase + i * RSTEP] : 0.0;
                        wv[i] = c * rcp_nr(ok ? part[i] : 1.0);
                    }
#pragma unroll
                    for (int j = 0; j < J; ++j) {
#pragma unroll
                        for (int i = 0; i < R; ++i) {
                            s[2 * j] = fma(wv[i], b[i][2 * j], s[2 * j]);
                            s[2 * j + 1] = fma(wv[i], b[i][2 * j + 1], s[2 * j + 1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < J; ++j)
                *reinterpret_cast<double2*>(spart + (size_t)rbase * KPAD + 2 * (kl + LK * j)) = make_double2(s[2 * j], s[2 * j + 1]);
            __syncthreads();
            // owners: CTA partial of the column sums (-> exchange slot of every CTA in the cluster)
            double* slot = xbuf + (size_t)(it & 1) * C * KPAD;
            double tot[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                double ss0 = 0.0, ss1 = 0.0, ss2 = 0.0, ss3 = 0.0;
                if (k < K) {
#pragma unroll
                    for (int q = 0; q < NP; q += 4) {
                        ss0 += spart[q * KPAD + k];
                        if (q + 1 < NP) ss1 += spart[(q + 1) * KPAD + k];
                        if (q + 2 < NP) ss2 += spart[(q + 2) * KPAD + k];
                        if (q + 3 < NP) ss3 += spart[(q + 3) * KPAD + k];
                    }
                }
                tot[u] = (ss0 + ss1) + (ss2 + ss3);
                if (C > 1 && k < K) {
                    for (uint32_t r = 0; r < C; ++r) st_cluster_f64(slot + crank * KPAD + k, r, tot[u]);
                }
            }
            if (C > 1) {
                cluster_arrive();
                cluster_wait();
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = gt + GT * u;
                    double ss = 0.0;
                    if (k < K) {
                        for (uint32_t r = 0; r < C; ++r) ss += slot[r * KPAD + k];
                    }
                    tot[u] = ss;
                }
            }
            // every CTA: same partials, same order -> bit-identical gamma / e / stop decision
            double gn[U], en[U];
            double dsum = 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                gn[u] = fma(er[u], tot[u], alr[u]);                       // :185
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
            {
                double* esn = es2 + (it & 1) * KPAD;                     // the buffer the NEXT trip reads
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = gt + GT * u;
                    if (k < K) esn[k] = en[u];
                }
            }
            // Stop rule (:187-190): sum_k |d gamma_k| <= tol K.  One thread's own share above tol K settles it
            // (the barrier carries that OR); only otherwise is the sum formed.
            if (__syncthreads_or(dsum > tolK)) {
                if (it >= p.max_iter) break;                              // :174
            } else {
                dsum = warp_sum(dsum);
                if (lane == 0) red[gw] = dsum;
                __syncthreads();
                dsum = 0.0;
#pragma unroll
                for (int w = 0; w < W; ++w) dsum += red[w];
                if (dsum <= tolK || it >= p.max_iter) break;              // :189-190 / :174
            }
#pragma unroll
            for (int u = 0; u < U; ++u) er[u] = en[u];
        }

        // ---- final pass: phi from the LAST e (buffer (it-1)&1; wv / part of the register rows are the last trip's)
        const double* es = es2 + ((it - 1) & 1) * KPAD;
        double lacc = 0.0;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = rbase + i * RSTEP;
            if (r < nR && kl == 0) lacc = fma(cnt[r], mwr[r] + log(part[i]), lacc);   // sum_n c_n logsumexp_n
        }
        {
            double e[2 * J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const double2 v = *reinterpret_cast<const double2*>(es + 2 * (kl + LK * j));
                e[2 * j] = v.x;
                e[2 * j + 1] = v.y;
            }
            double* rowp = tile + (size_t)rbase * ST + 2 * kl;
            for (int r0 = gw * LN; r0 < nS; r0 += RSTEP, rowp += (size_t)RSTEP * ST) {
                const int r = r0 + nl;
                const bool ok = r < nS;
                double bb[2 * J];
                const double pt = row_dot<LK, J>(rowp, e, bb, ok ? min(J, (KP2 - kl + LK - 1) / LK) : 0);
                const double c = cnt[nR + r];
                const double w = ok ? c * rcp_nr(pt) : 0.0;
                if (ok && kl == 0) lacc = fma(c, mwr[nR + r] + log(pt), lacc);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (ok && kl + LK * j < KP2)
                        *reinterpret_cast<double2*>(rowp + 2 * LK * j) =
                            make_double2(w * bb[2 * j] * e[2 * j], w * bb[2 * j + 1] * e[2 * j + 1]);   // :207
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
        __syncthreads();     // phi of the shared-memory rows complete
        for (int r = gt; r < nS; r += GT)
            bulk_red_add_f64(p.phi_ss + (size_t)rid[nR + r] * KP, tile + (size_t)r * ST, (uint32_t)KP * 8u);
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
        // the register rows leave through the same region once the reduce-adds above have read it
        bulk_wait_read0();
        __syncthreads();
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = rbase + i * RSTEP;
            if (r < nR) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (kl + LK * j < KP2) {
                        const double2 ev = *reinterpret_cast<const double2*>(es + 2 * (kl + LK * j));
                        *reinterpret_cast<double2*>(tile + (size_t)r * ST + 2 * (kl + LK * j)) =
                            make_double2(wv[i] * b[i][2 * j] * ev.x, wv[i] * b[i][2 * j + 1] * ev.y);   // :207
                    }
                }
            }
        }
        fence_async_smem();
        __syncthreads();
        for (int r = gt; r < nR; r += GT)
            bulk_red_add_f64(p.phi_ss + (size_t)rid[r] * KP, tile + (size_t)r * ST, (uint32_t)KP * 8u);
        bulk_commit();
    }
    bulk_wait_read0();
    if (C > 1) {
        cluster_arrive();   // nobody leaves while a peer may still address its shared memory
        cluster_wait();
    }
}

}  // namespace pylda
