// Per-document VB E-step kernel, second generation (sm_100a) -- resident-tile path.
//
// Same mathematics as estep_kernel.cuh (reference variational_bayes.py:159-207 in product
// form); what changed is the execution shape, driven by the first ncu captures
// (profiles/r1a_*: 3600 warp instructions per document-trip, 19.5 % of them DFMA, 8 warps/SM,
// 42 % of stall samples in fixed-latency dependency waits):
//   * W (warps per document group) is a template parameter: single-warp groups (W = 1) run the
//     whole fixed point with __syncwarp only, multi-warp groups use two named barriers per trip;
//   * the exponent shift c is dropped (e_k = exp(psi(gamma_k)) cannot overflow and underflows
//     exactly where the reference's exp() does), so no per-trip max reduction / log;
//   * gamma_k, e_k and alpha_k of the owner thread stay in registers across trips;
//   * the next trip's e is computed speculatively while the convergence sum is reduced;
//   * column partial sums go through shared memory as a reduce-scatter (J STS.128 per lane,
//     owners read W*LN values) instead of 2J shuffle butterflies per lane;
//   * Newton reciprocal (MUFU.RCP64H + 2 steps) instead of IEEE division, no slow-path branch;
//   * tile loads are unpredicated: the last topic pairs of a row may over-read into the next row
//     (finite data) and are multiplied by e = 0;
//   * the queue index of the next document is prefetched.
#pragma once
#include "estep_kernel.cuh"
#include "estep_narrow.cuh"

namespace pylda {

template <int W>
__device__ __forceinline__ void gsync(int g) {
    if (W == 1) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(W * 32) : "memory");
    }
}

// npairs: pairs of this lane that lie inside the row (kl + LK*j < KP/2).  The trip loops pass J (no
// predicate: the over-read of the next row is multiplied by e = 0 and nobody writes the tile then); the
// final passes, which overwrite tile rows with phi while other warps still read their own rows, pass the
// true count so that no lane ever reads a row it does not own (compute-sanitizer racecheck clean).
template <int LK, int J>
__device__ __forceinline__ double row_dot(const double* rowp, const double (&e)[2 * J], double (&b)[2 * J],
                                          int npairs = J) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
        double2 v = make_double2(0.0, 0.0);
        if (j < npairs) v = *reinterpret_cast<const double2*>(rowp + 2 * LK * j);
        b[2 * j] = v.x;
        b[2 * j + 1] = v.y;
    }
    // four independent accumulation chains
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        if (j & 1) {
            a2 = fma(b[2 * j], e[2 * j], a2);
            a3 = fma(b[2 * j + 1], e[2 * j + 1], a3);
        } else {
            a0 = fma(b[2 * j], e[2 * j], a0);
            a1 = fma(b[2 * j + 1], e[2 * j + 1], a1);
        }
    }
    double part = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 1; o < LK; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    return part;
}

// RR row groups (RR*LN rows per warp) of one trip: norm_n = B[n,:].e, w_n = c_n / norm_n,
// s += w_n B[n,:].  The RR rows of a lane are independent dependency chains (ILP).
template <int LK, int J, int RR>
__device__ __forceinline__ void rows_accum(const double* rowp, size_t gstride, const double* cntp, int cstride,
                                           const double (&e)[2 * J], double (&s)[2 * J]) {
    double b[RR][2 * J];
    double part[RR];
#pragma unroll
    for (int i = 0; i < RR; ++i) part[i] = row_dot<LK, J>(rowp + i * gstride, e, b[i]);
    double w[RR];
#pragma unroll
    for (int i = 0; i < RR; ++i) w[i] = cntp[i * cstride] * rcp_nr(part[i]);
#pragma unroll
    for (int i = 0; i < RR; ++i) {
#pragma unroll
        for (int c = 0; c < 2 * J; ++c) s[c] = fma(w[i], b[i][c], s[c]);
    }
}

// V = 0: CTAs of up to 8 warps (255 registers available, two row groups in flight per lane);
// V = 1: CTAs of up to 16 warps (128 registers, one row group per lane) -- more documents or more
//        warps per document in flight per SM, to overlap one document's serial owner phase with
//        another's row phase.
template <int LK, int J, int W, int V>
__global__ void __launch_bounds__(V == 1 ? 512 : 256) estep_v2(const EParams p) {
    constexpr int LN = 32 / LK;
    constexpr int KPAD = 2 * LK * J;
    constexpr int GT = 32 * W;
    constexpr int U = (KPAD + GT - 1) / GT;        // topics per owner thread
    constexpr bool BFLY = (W >= 4);                // reduce over the LN row-lanes by shuffles first
    constexpr int NP = BFLY ? W : W * LN;          // partial rows in spart
    constexpr int RR = (V == 0 && 2 * J <= 16) ? 2 : 1;   // row groups per trip-loop iteration (ILP vs registers)
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int g = tid / GT;
    const int gt = tid - g * GT;
    const int gw = gt >> 5;
    const int lane = tid & 31;
    const int kl = lane % LK;
    const int nl = lane / LK;
    const int K = p.K, KP = p.KP, ST = p.ST;
    const int KP2 = KP >> 1;

    unsigned char* gs = smem_raw + (size_t)g * p.group_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(gs);
    int* cur = reinterpret_cast<int*>(gs + 8);
    double* es = reinterpret_cast<double*>(gs + 16);
    double* spart = reinterpret_cast<double*>(gs + p.off_spart);
    double* red = reinterpret_cast<double*>(gs + p.off_red);
    double* cnt = reinterpret_cast<double*>(gs + p.off_cnt);
    double* mwr = reinterpret_cast<double*>(gs + p.off_mwr);
    int* rid = reinterpret_cast<int*>(gs + p.off_rid);
    double* tile = reinterpret_cast<double*>(gs + p.off_tile);

    // one-time: finite contents everywhere (stale / over-read rows are multiplied by 0), barrier
    for (int i = 16 + gt * 8; i < p.group_bytes; i += GT * 8) *reinterpret_cast<double*>(gs + i) = 0.0;
    if (gt == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // owner threads: topic k = gt + GT*u lives in registers for the whole kernel
    double alr[U], gamr[U], er[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int k = gt + GT * u;
        alr[u] = (k < K) ? p.alpha[k] : 1.0;
        gamr[u] = 1.0;
        er[u] = 0.0;
    }

    const bool warp_owns = gw * 32 < K;   // warps whose lanes own no topic skip the exp(psi) work
    uint32_t parity = 0;
    int nxt = 0;
    if (gt == 0) nxt = atomicAdd(p.counter, 1);

    while (true) {
        // ---- next document of this class; the previous one is fully retired ----------------
        bulk_wait_read0();   // this thread's reduce-adds have finished reading the tile
        int idx;
        if (W == 1) {
            idx = __shfl_sync(0xffffffffu, nxt, 0);
        } else {
            if (gt == 0) *cur = nxt;
            gsync<W>(g);
            idx = *cur;
        }
        if (idx >= p.ndocs) break;
        if (gt == 0) nxt = atomicAdd(p.counter, 1);   // prefetch: consumed one document later
        const int d = p.order[idx];
        const long long base = p.row_ptr[d];
        const int n = (int)(p.row_ptr[d + 1] - base);

        // ---- stage ids / counts / m_w and the B tile (bulk-async row copies) ------------------
        // Rows n .. npad-1 (npad = n rounded up to LN) are copies of row 0 with count 0: every row
        // group is then uniform (positive norm, zero weight) and the trip loop needs no row predicate.
        const int npad = (n + LN - 1) / LN * LN;
        const int NG = npad / LN;
        if (gt == 0) mbar_expect_tx(mbar, (uint32_t)npad * (uint32_t)KP * 8u);
        int csum = 0;
        for (int r = gt; r < npad; r += GT) {
            const bool real = r < n;
            const int id = p.ids[base + (real ? r : 0)];
            const int c = real ? p.cts[base + r] : 0;
            rid[r] = id;
            cnt[r] = (double)c;
            mwr[r] = p.mw[id];
            csum += c;
            bulk_g2s(tile + (size_t)r * ST, p.Bt + (size_t)id * KP, (uint32_t)KP * 8u, mbar);
        }
        csum = __reduce_add_sync(0xffffffffu, csum);
        double Nd = (double)csum;
        if (W > 1) {
            if (lane == 0) red[gw] = Nd;
            gsync<W>(g);
            Nd = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) Nd += red[w];
        }
        // gamma0 = alpha + N_d / K                              (variational_bayes.py:165)
        const double g0 = Nd / (double)K;
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
        gsync<W>(g);

        // ---- fixed-point trips                                 (variational_bayes.py:174-190)
        double e[2 * J];
        int it = 0;
        const double tolK = p.tol * (double)K;
        bool parked = false;
        const bool long_park = W > 1 && n > 192 && n <= p.park_long;
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
                // warp gw owns row groups q = gw, gw + W, ...; RR groups per iteration, then the tail
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
            // column sums: reduce-scatter through shared memory
            if (BFLY) {
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
            } else {
#pragma unroll
                for (int j = 0; j < J; ++j)
                    *reinterpret_cast<double2*>(spart + (gw * LN + nl) * KPAD + 2 * (kl + LK * j)) =
                        make_double2(s[2 * j], s[2 * j + 1]);
            }
            gsync<W>(g);
            // owners: gamma update (:185), |d gamma| (:187), speculative e for the next trip
            double gn[U], en[U];
            double dsum = 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                double ss0 = 0.0, ss1 = 0.0;
                if (k < K) {
#pragma unroll
                    for (int q = 0; q < NP; q += 2) {
                        ss0 += spart[q * KPAD + k];
                        if (q + 1 < NP) ss1 += spart[(q + 1) * KPAD + k];
                    }
                }
                gn[u] = fma(er[u], ss0 + ss1, alr[u]);
                if (k < K) dsum += fabs(gn[u] - gamr[u]);
            }
            if (warp_owns) {
#pragma unroll
                for (int u = 0; u < U; ++u) en[u] = exp_digamma(gn[u]);
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) en[u] = 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) gamr[u] = gn[u];                 // :188
            ++it;
            dsum = warp_sum(dsum);
            unsigned bal[U];
            if (W > 1) {
                if (lane == 0) red[gw] = dsum;
                if (long_park) {
                    // live topics of this warp's owners, counted with the same barrier as |d gamma|
                    int mine = 0;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int k = gt + GT * u;
                        bal[u] = __ballot_sync(0xffffffffu, k < K && gn[u] != alr[u]);
                        mine += __popc(bal[u]);
                    }
                    if (lane == 0) red[W + gw] = (double)mine;   // (red[W ..) belongs to the ELBO exchange of the final pass)
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = gt + GT * u;
                    if (k < K) es[k] = en[u];
                }
                gsync<W>(g);
                dsum = 0.0;
#pragma unroll
                for (int w = 0; w < W; ++w) dsum += red[w];
            }
            if (dsum <= tolK || it >= p.max_iter) break;                 // :189-190 / :174
            if (long_park) {
                // At most 32 topics alive (gamma_k != alpha_k) and enough trips left: estep_longc finishes the
                // document on a compact tile (same hand-over as in estep_stream below).
                int nlive = 0, rank = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    const int c = (int)red[W + w];
                    nlive += c;
                    if (w < gw) rank += c;
                }
                if (nlive >= 1 && nlive <= 32 && it + PARK_LONG_MIN_TRIPS <= p.max_iter) {
                    int* rec = p.park_rec + (size_t)d * PARK_REC;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int k = gt + GT * u;
                        if (k < K) p.gamma[(size_t)d * K + k] = gamr[u];       // final for every dead topic
                        if ((bal[u] >> lane) & 1u) {
                            const int slot = rank + __popc(bal[u] & ((1u << lane) - 1u));
                            rec[2 + slot] = k;
                            p.park_gam[(size_t)d * PARK_GAM + slot] = gamr[u];
                        }
                        rank += __popc(bal[u]);
                    }
                    if (gt == 0) {
                        rec[0] = it;
                        rec[1] = nlive;
                        const int li = park_list_index(nlive, n);
                        const int slot = atomicAdd(p.park_counts + li, 1);
                        p.park_lists[(size_t)li * p.park_cap + slot] = d;
                    }
                    parked = true;
                    break;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) er[u] = en[u];
            if (W == 1) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = gt + GT * u;
                    if (k < K) es[k] = en[u];
                }
                __syncwarp();
            }
        }

        if (parked) continue;
        // ---- final pass: phi from the LAST e (registers e[], owners' er[]) -------------------
        double lacc = 0.0;
        {
            double* rowp = tile + (size_t)(gw * LN + nl) * ST + 2 * kl;
            for (int r0 = gw * LN; r0 < n; r0 += W * LN, rowp += (size_t)W * LN * ST) {
                const int r = r0 + nl;
                const bool ok = r < n;
                double b[2 * J];
                const double part = row_dot<LK, J>(rowp, e, b, ok ? min(J, (KP2 - kl + LK - 1) / LK) : 0);
                const double c = cnt[r];
                const double w = ok ? c * rcp_nr(part) : 0.0;
                if (ok && kl == 0) lacc = fma(c, mwr[r] + log(part), lacc);   // sum_n c_n logsumexp_n
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (ok && kl + LK * j < KP2)
                        *reinterpret_cast<double2*>(rowp + 2 * LK * j) =
                            make_double2(w * b[2 * j] * e[2 * j], w * b[2 * j + 1] * e[2 * j + 1]);   // c_n phi_nk (:207)
                }
            }
        }
        fence_async_smem();   // generic-proxy writes of phi -> visible to the bulk-async engine
        // ---- per-document ELBO pieces and gamma write-back ------------------------------------
        double t1 = lacc, sg = 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = gt + GT * u;
            if (k < K) {
                const double gk = gamr[u];
                const double ek = er[u];
                const double dk = gk - alr[u];
                t1 += lgamma(gk);                                            // :197
                if (ek > 0.0 && dk != 0.0) t1 -= log(ek) * dk;               // - sum_k psi_k sum_n c_n phi_nk
                sg += gk;
                p.gamma[(size_t)d * K + k] = gk;                             // :212 / :216
            }
        }
        t1 = warp_sum(t1);
        sg = warp_sum(sg);
        if (W > 1) {
            if (lane == 0) {
                red[W + 2 * gw] = t1;
                red[W + 2 * gw + 1] = sg;
            }
        }
        gsync<W>(g);   // all phi rows written (and red[] complete)
        for (int r = gt; r < n; r += GT)
            bulk_red_add_f64(p.phi_ss + (size_t)rid[r] * KP, tile + (size_t)r * ST, (uint32_t)KP * 8u);
        bulk_commit();
        if (gt == 0) {
            if (W > 1) {
                t1 = 0.0;
                sg = 0.0;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    t1 += red[W + 2 * w];
                    sg += red[W + 2 * w + 1];
                }
            }
            p.docterm[d] = t1 - lgamma(sg);                                  // - lgamma(sum_k gamma_k), :197
            p.iters[d] = it;
        }
    }
}

}  // namespace pylda

namespace pylda {

// ---- streaming variant: documents too long for any resident class -----------------------------
// One CTA of 8 warps per document, two CTAs per SM.  The B rows are NOT staged: every trip re-reads
// them from L2 (the V x KP table is 80 MB at the headline config, L2 is 126 MB) with unpredicated
// LDG.128, two row groups in flight per lane; only term ids and counts live in shared memory.
// Same trip structure as estep_v2 with W = 8; phi leaves by red.global.add.f64.
template <int LK, int J, int RR>
__device__ __forceinline__ void rows_accum_global(const double* __restrict__ Bt, int KP, int kl, const int* ridp,
                                                  const double* cntp, int gstride_rows, const double (&e)[2 * J],
                                                  double (&s)[2 * J], int npairs = J) {
    double b[RR][2 * J];
    double part[RR];
#pragma unroll
    for (int i = 0; i < RR; ++i) {
        const double* rowp = Bt + (size_t)ridp[i * gstride_rows] * KP + 2 * kl;
        part[i] = row_dot<LK, J>(rowp, e, b[i], npairs);
    }
#pragma unroll
    for (int i = 0; i < RR; ++i) {
        const double w = cntp[i * gstride_rows] * rcp_nr(part[i]);
#pragma unroll
        for (int c = 0; c < 2 * J; ++c) s[c] = fma(w, b[i][c], s[c]);
    }
}

// MODE 0: lean instantiation for classes whose documents can neither be handed over (all longer than 192 terms and
//         the compact stage for long documents off) nor need chunked staging;
// MODE 2: hand-over to the narrow stages / estep_longc, no chunked staging;
// MODE 1: both.  The extra live state of each feature costs the 128-register kernel spills in its trip loop, hence
// one instantiation per combination in use.
template <int LK, int J, int MODE>
__global__ void __launch_bounds__(256, 2) estep_stream(const EParams p) {
    constexpr bool CHUNK = (MODE == 1);
    constexpr bool PARK = (MODE >= 1);
    constexpr int W = 8;
    constexpr int LN = 32 / LK;
    constexpr int KPAD = 2 * LK * J;
    constexpr int GT = 256;
    constexpr int U = (KPAD + GT - 1) / GT;
    constexpr int RR = (2 * J <= 16) ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int gt = threadIdx.x;
    const int gw = gt >> 5;
    const int lane = gt & 31;
    const int kl = lane % LK;
    const int nl = lane / LK;
    const int K = p.K, KP = p.KP;
    const int KP2 = KP >> 1;

    unsigned char* gs = smem_raw;
    int* cur = reinterpret_cast<int*>(gs + 8);
    double* es = reinterpret_cast<double*>(gs + 16);
    double* spart = reinterpret_cast<double*>(gs + p.off_spart);     // [W][KPAD]
    double* red = reinterpret_cast<double*>(gs + p.off_red);
    double* cnt = reinterpret_cast<double*>(gs + p.off_cnt);
    int* rid = reinterpret_cast<int*>(gs + p.off_rid);

    for (int i = 16 + gt * 8; i < p.group_bytes; i += GT * 8) *reinterpret_cast<double*>(gs + i) = 0.0;
    __syncthreads();

    double alr[U], gamr[U], er[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int k = gt + GT * u;
        alr[u] = (k < K) ? p.alpha[k] : 1.0;
        gamr[u] = 1.0;
        er[u] = 0.0;
    }
    const bool warp_owns = gw * 32 < K;
    int nxt = 0;
    if (gt == 0) nxt = atomicAdd(p.counter, 1);

    while (true) {
        if (gt == 0) *cur = nxt;
        __syncthreads();
        const int idx = *cur;
        if (idx >= p.ndocs) break;
        if (gt == 0) nxt = atomicAdd(p.counter, 1);
        const int d = p.order[idx];
        const long long base = p.row_ptr[d];
        const int n = (int)(p.row_ptr[d + 1] - base);
        const int npad = (n + LN - 1) / LN * LN;
        // ids and counts live in shared memory, p.nmax rows at a time: a document longer than that (thousands of
        // terms) is walked in chunks that are re-staged every trip -- 12 bytes per row against the 8 K of its B row
        const int cap = p.nmax;                             // multiple of W * LN
        const int nch = CHUNK ? (npad + cap - 1) / cap : 1;
        // (a macro, not a lambda: by-reference captures put the captured variables on the stack)
#define PYLDA_STAGE_CHUNK(CH)                                                                          \
        {                                                                                              \
            const int r0_ = (CH) * cap;                                                                \
            rows_c = min(cap, npad - r0_);                                                             \
            for (int r = gt; r < rows_c; r += GT) {                                                    \
                const bool real = r0_ + r < n;                                                         \
                rid[r] = p.ids[base + (real ? r0_ + r : 0)];     /* pad rows: a valid row, weight 0 */ \
                cnt[r] = real ? (double)p.cts[base + r0_ + r] : 0.0;                                   \
            }                                                                                          \
        }
        int csum = 0;
        for (int r = gt; r < n; r += GT) csum += p.cts[base + r];
        int rows_c;
        PYLDA_STAGE_CHUNK(0)
        csum = __reduce_add_sync(0xffffffffu, csum);
        if (lane == 0) red[gw] = (double)csum;
        __syncthreads();
        double Nd = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) Nd += red[w];
        const double g0 = Nd / (double)K;                   // gamma0 = alpha + N_d / K   (:165)
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
        __syncthreads();

        double e[2 * J];
        int it = 0;
        const double tolK = p.tol * (double)K;
        // hand-over threshold of this document: 32 live topics for n <= 96, 16 for n <= 192; documents longer than
        // that: estep_longc at 32 live topics.  (Recomputed where it is used: one value less to keep across the row loop.)
#define PYLDA_PARK_THR() ((p.park_nc >= 16 && n <= 96) ? 32 : (p.park_nc > 0 && n <= 192) ? p.park_nc \
                                                            : (n > 192 && n <= p.park_long) ? 32 : 0)
        bool parked = false;
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
            for (int ch = 0; ch < nch; ++ch) {
                if (nch > 1) {
                    __syncthreads();
                    PYLDA_STAGE_CHUNK(ch)
                    __syncthreads();
                }
                const int NG = rows_c / LN;
                // the warp's row groups gw, gw + W, ...: forwards on even trips, backwards on odd ones, so that the
                // rows read last are read first again and the L1 cache (a fraction of the tile) is not swept
                // cyclically -- every trip gets ~L1/tile of its rows from L1 instead of none
                const int* ridp = rid + gw * LN + nl;
                const double* cntp = cnt + gw * LN + nl;
                const int M = (NG - gw + W - 1) / W;
                constexpr int GS = W * LN;
                if (!(it & 1) || !p.compact) {
                    int m = 0;
                    for (; m + RR <= M; m += RR)
                        rows_accum_global<LK, J, RR>(p.Bt, KP, kl, ridp + m * GS, cntp + m * GS, GS, e, s);
                    if (RR > 1) {
                        for (; m < M; ++m)
                            rows_accum_global<LK, J, 1>(p.Bt, KP, kl, ridp + m * GS, cntp + m * GS, GS, e, s);
                    }
                } else {
                    int m = M;
                    for (; m >= RR; m -= RR)
                        rows_accum_global<LK, J, RR>(p.Bt, KP, kl, ridp + (m - RR) * GS, cntp + (m - RR) * GS, GS, e, s);
                    if (RR > 1) {
                        for (; m > 0; --m)
                            rows_accum_global<LK, J, 1>(p.Bt, KP, kl, ridp + (m - 1) * GS, cntp + (m - 1) * GS, GS, e, s);
                    }
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
            double gn[U], en[U];
            double dsum = 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                double ss0 = 0.0, ss1 = 0.0;
                if (k < K) {
#pragma unroll
                    for (int q = 0; q < W; q += 2) {
                        ss0 += spart[q * KPAD + k];
                        ss1 += spart[(q + 1) * KPAD + k];
                    }
                }
                gn[u] = fma(er[u], ss0 + ss1, alr[u]);                    // :185
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
            // live topics (gamma_k != alpha_k) of this warp's owners: counted with the same barrier as |d gamma|
            if (PARK && PYLDA_PARK_THR() > 0) {
                int mine = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = gt + GT * u;
                    mine += __popc(__ballot_sync(0xffffffffu, k < K && gn[u] != alr[u]));
                }
                if (lane == 0) red[W + gw] = (double)mine;      // (red[W ..) belongs to the ELBO exchange of the final pass)
            }
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
            if (PARK && PYLDA_PARK_THR() > 0) {
                const int park_thr = PYLDA_PARK_THR();
                // Few enough topics alive: the narrow stages (estep_narrow.cuh; documents of up to 192 terms) or the
                // compact stage for long documents (estep_longc.cuh) finish the document.  This kernel has no compact
                // stage of its own, so it hands over at 32 live topics already.  A long document is only worth the
                // hand-over (one gather of its live columns) while enough trips are left.
                int nlive = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) nlive += (int)red[W + w];
                if (nlive >= 1 && nlive <= park_thr && (n <= 192 || it + PARK_LONG_MIN_TRIPS <= p.max_iter)) {
                    int* rec = p.park_rec + (size_t)d * PARK_REC;
                    int rank = 0;
                    for (int w = 0; w < gw; ++w) rank += (int)red[W + w];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int k = gt + GT * u;
                        const unsigned bal = __ballot_sync(0xffffffffu, k < K && gamr[u] != alr[u]);   // gamr = the new gamma
                        if (k < K) p.gamma[(size_t)d * K + k] = gamr[u];       // final for every dead topic
                        if ((bal >> lane) & 1u) {
                            const int slot = rank + __popc(bal & ((1u << lane) - 1u));
                            rec[2 + slot] = k;
                            p.park_gam[(size_t)d * PARK_GAM + slot] = gamr[u];
                        }
                        rank += __popc(bal);
                    }
                    if (gt == 0) {
                        rec[0] = it;
                        rec[1] = nlive;
                        const int li = park_list_index(nlive, n);
                        const int slot = atomicAdd(p.park_counts + li, 1);
                        p.park_lists[(size_t)li * p.park_cap + slot] = d;
                    }
                    parked = true;
                    break;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) er[u] = en[u];
        }
        if (parked) continue;

        // ---- final pass: phi from the LAST e, scattered with red.global.add.f64 -----------------
        double lacc = 0.0;
        for (int ch = 0; ch < nch; ++ch) {
            if (nch > 1) {
                __syncthreads();
                PYLDA_STAGE_CHUNK(ch)
                __syncthreads();
            }
            const int rbase = ch * cap;
            for (int r0 = gw * LN; r0 < rows_c; r0 += W * LN) {
                const int r = r0 + nl;
                const bool ok = rbase + r < n;
                const int id = rid[r];
                double b[2 * J];
                const double part = row_dot<LK, J>(p.Bt + (size_t)id * KP + 2 * kl, e, b);
                const double c = cnt[r];
                const double w = ok ? c * rcp_nr(part) : 0.0;
                if (ok && kl == 0) lacc = fma(c, p.mw[id] + log(part), lacc);
                double* dst = p.phi_ss + (size_t)id * KP + 2 * kl;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (ok && kl + LK * j < KP2) {
                        atomicAdd(dst + 2 * LK * j, w * b[2 * j] * e[2 * j]);                       // :207
                        if (2 * (kl + LK * j) + 1 < K) atomicAdd(dst + 2 * LK * j + 1, w * b[2 * j + 1] * e[2 * j + 1]);
                    }
                }
            }
        }
        double t1 = lacc, sg = 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = gt + GT * u;
            if (k < K) {
                const double gk = gamr[u];
                const double ek = er[u];
                const double dk = gk - alr[u];
                t1 += lgamma(gk);                                            // :197
                if (ek > 0.0 && dk != 0.0) t1 -= log(ek) * dk;
                sg += gk;
                p.gamma[(size_t)d * K + k] = gk;                             // :212 / :216
            }
        }
        t1 = warp_sum(t1);
        sg = warp_sum(sg);
        if (lane == 0) {
            red[W + 2 * gw] = t1;
            red[W + 2 * gw + 1] = sg;
        }
        __syncthreads();
        if (gt == 0) {
            t1 = 0.0;
            sg = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                t1 += red[W + 2 * w];
                sg += red[W + 2 * w + 1];
            }
            p.docterm[d] = t1 - lgamma(sg);                                  // - lgamma(sum_k gamma_k), :197
            p.iters[d] = it;
        }
    }
}

#undef PYLDA_STAGE_CHUNK
#undef PYLDA_PARK_THR

}  // namespace pylda
