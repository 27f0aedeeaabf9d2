// Per-document VB E-step kernel, compact stage for LONG documents (sm_100a).
//
// Same mathematics as the other generations (reference variational_bayes.py:174-207 in product form).  Why it
// exists: a document of hundreds or thousands of terms loses its topics later than a short one, but it does lose
// them -- at the cold state of the headline config a 200..270-term document has <= 32 topics alive from trip ~31,
// and in the warm state of a training run (EM iteration >= 2) a 1000-term document is down to <= 32 at trip ~20 and
// <= 16 at trip ~30 (scripts/sim_live.py, scripts/tune_lda.py).  estep_stream re-reads the full 8 K bytes of every
// row from L2 on every trip, estep_v2 the full shared-memory tile.  Both therefore PARK a document longer than 192
// terms once at most 32 topics are alive (list 9: gamma -- final for every dead topic --, the trip count and the live
// columns, as for the narrow stages of estep_narrow.cuh), and this kernel finishes it:
//
//   one CTA of 8 warps per document, two CTAs per SM.  The CTA gathers the live columns of the document's rows once
//   into a compact tile of NC = 32 doubles per row -- in shared memory when the document fits (<= smem_rows rows),
//   else in a per-CTA scratch area in global memory (L2-resident: 256 bytes per row instead of 8 K) -- and runs the
//   remaining trips on it: 4 lanes per row (8 columns each), 8 rows per warp step, the same owner phase as the
//   narrow stages (32 owner threads, one live topic each).
//
// phi and the validation of the elimination are those of the narrow stages: live columns scattered with (e - e_dead),
// row weights into wsum for k_dead_phi; a document whose weights sum above chk_bound gets the full-width check.
#pragma once
#include "estep_narrow.cuh"

namespace pylda {

struct LParams {
    NParams n;                 // the narrow stages' parameters (list = list 9)
    double* scratch_tile;      // per CTA: scratch_rows * 32 doubles
    double* scratch_cnt;       // per CTA: scratch_rows doubles (c_n)
    int* scratch_ids;          // per CTA: scratch_rows ints (term ids)
    int scratch_rows;          // capacity of one CTA's scratch (multiple of 64)
    int smem_rows;             // rows of the shared-memory tile (multiple of 64)
    int mix;                   // every mix-th claim takes the next document from the front of the list, the others from
                               // its back (1: plain order)
};

// L2 eviction priorities (createpolicy / ld.global.L2::cache_hint): the scratch tiles are re-read every trip and
// should stay in L2; the rows of the (V, KP) table gathered into them, and the tile on its last pass, should not
// push them out (measured without the hints: 40 % of the tile bytes came from DRAM, profiles/r2g_longc_*).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double2 ldg_hint_v2(const double* ptr, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ double ldg_hint(const double* ptr, uint64_t pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg_hint(double* ptr, double v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(ptr), "d"(v), "l"(pol) : "memory");
}

// RR row groups of one trip on the compact tile: norm_n = T[n,:].e, w_n = c_n / norm_n, s += w_n T[n,:].  A lane
// holds 8 columns of a row; the RR rows of a lane are independent chains (all their loads are in flight together).
// GLB: the tile is in the global scratch (loads carry the L2 policy `pol`), else in shared memory.
template <int NC, int RR, bool GLB>
__device__ __forceinline__ void longc_rows(const double* rowp, const double* cntp, const double (&e)[8], double (&s)[8],
                                           uint64_t pol, const int (&po)[4]) {
    constexpr int CPL = 8, LK = NC / CPL, GS = 8 * (32 / LK);
    double b[RR][CPL], c[RR], part[RR];
#pragma unroll
    for (int q = 0; q < RR; ++q) {
#pragma unroll
        for (int i = 0; i < CPL; i += 2) {
            const double2 v = GLB ? ldg_hint_v2(rowp + (size_t)q * GS * NC + po[i >> 1], pol)
                                  : *reinterpret_cast<const double2*>(rowp + (size_t)q * GS * NC + po[i >> 1]);
            b[q][i] = v.x;
            b[q][i + 1] = v.y;
        }
        c[q] = cntp[q * GS];
    }
#pragma unroll
    for (int q = 0; q < RR; ++q) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i = 0; i < CPL; i += 2) {
            a0 = fma(b[q][i], e[i], a0);
            a1 = fma(b[q][i + 1], e[i + 1], a1);
        }
        part[q] = a0 + a1;
    }
#pragma unroll
    for (int o = 1; o < LK; o <<= 1) {
#pragma unroll
        for (int q = 0; q < RR; ++q) part[q] += __shfl_xor_sync(0xffffffffu, part[q], o);
    }
#pragma unroll
    for (int q = 0; q < RR; ++q) {
        const double w = c[q] * rcp_nr(c[q] > 0.0 ? part[q] : 1.0);
#pragma unroll
        for (int i = 0; i < CPL; ++i) s[i] = fma(w, b[q][i], s[i]);
    }
}

// RR row groups in flight per lane; MINB CTAs per SM (register budget 65536 / (256 MINB))
template <int NC, int RR, int MINB>
__global__ void __launch_bounds__(256, MINB) estep_longc(const LParams lp) {
    constexpr int W = 8;
    constexpr int CPL = 8;                 // columns per lane
    constexpr int LK = NC / CPL;           // lanes per row
    static_assert(NC == 32, "the bank swizzle below is written for 16 column pairs per row");
    constexpr int LN = 32 / LK;            // rows per warp step
    constexpr int GS = W * LN;             // rows per CTA step
    const NParams& p = lp.n;
    extern __shared__ __align__(16) double lsm[];
    double* es = lsm;                      // [NC] e of the live slots
    double* spart = es + NC;               // [W][NC] column partials
    double* red = spart + W * NC;          // [4 * W] reductions
    int* ctl = reinterpret_cast<int*>(red + 4 * W);    // [0] document index, [1] stop flag, [2] came-back flag
    double* sk = red + 4 * W + 2;          // [K] column sums of the validation pass
    double* tile_s = sk + ((p.K + 1) & ~1);

    const int gt = threadIdx.x, gw = gt >> 5, lane = gt & 31;
    const int kl = lane % LK, nl = lane / LK;
    const int K = p.K, KP = p.KP;
    const double tolK = p.tol * (double)K;
    const int ndocs = *p.count;
    // Column layout.  A lane owns the column pairs kl + LK t, t = 0..3 (registers 2t, 2t+1 <-> columns 2 (kl + LK t) + h):
    // for one t the LK lanes of a row read 64 contiguous bytes.  In a row of odd index the two 64-byte halves of
    // every 128 bytes are swapped, so that the two rows of a quarter warp hit disjoint shared-memory banks (rows are
    // 256 bytes apart: without the swap every LDS.128 was a 4-way bank conflict, profiles/r2g_longc_cold_raw.txt).
    auto lcol = [&](int i) { return 2 * (kl + LK * (i >> 1)) + (i & 1); };                 // register -> logical column
    auto pcol = [&](int j, int r) { return (((j >> 1) ^ ((r & 1) << 2)) << 1) | (j & 1); };   // logical -> position in row r
    int po[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) po[t] = 2 * (kl + LK * (t ^ (nl & 1)));
    const bool owner = gt < NC;
    const uint64_t keep = l2_policy_evict_last(), once = l2_policy_evict_first();

    double* tile_g = lp.scratch_tile + (size_t)blockIdx.x * lp.scratch_rows * NC;
    double* cs = lp.scratch_cnt + (size_t)blockIdx.x * lp.scratch_rows;
    int* is = lp.scratch_ids + (size_t)blockIdx.x * lp.scratch_rows;

    while (true) {
        __syncthreads();                                   // the previous document is fully retired
        if (gt == 0) ctl[0] = atomicAdd(p.head, 1);
        __syncthreads();
        const int claim = ctl[0];
        if (claim >= ndocs) break;
        // The list is roughly in descending length (estep_stream works longest-first, the 193..272-term documents of
        // estep_v2 come last).  Taken in that order, all CTAs would hold their longest tiles at the same time -- 2-3x
        // the L2 -- so one claim in `mix` takes from the front and the others from the back: a permutation of the list.
        int idx = claim;
        if (lp.mix > 1) {
            const int q = claim / lp.mix;
            idx = (claim % lp.mix == 0) ? q : ndocs - 1 - (claim - q - 1);
        }
        const int d = p.list[idx];
        const long long base = p.row_ptr[d];
        const int n = (int)(p.row_ptr[d + 1] - base);
        const int npad = (n + GS - 1) / GS * GS;
        const int* rec = p.rec + (size_t)d * PARK_REC;
        int it = rec[0];
        const int nlive = rec[1];
        const bool glb = npad > lp.smem_rows;
        double* T = glb ? tile_g : tile_s;

        // ---- term ids and counts (coalesced), then the compact tile: the live columns of every row, gathered from
        //      the (V, KP) table once; four rows per warp in flight ----
        for (int r = gt; r < npad; r += 256) {
            const bool ok = r < n;
            is[r] = ok ? p.ids[base + r] : 0;
            cs[r] = ok ? (double)p.cts[base + r] : 0.0;
        }
        __syncthreads();
        for (int j = lane; j < NC; j += 32) {
            const int col = (j < nlive) ? rec[2 + j] : -1;
            for (int r = gw; r < npad; r += 4 * W) {
                double v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = r + q * W;
                    v[q] = (rr < n && col >= 0) ? ldg_hint(p.Bt + (size_t)is[rr] * KP + col, once) : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = r + q * W;
                    if (rr < npad) {
                        const int pj = pcol(j, rr);
                        if (glb) stg_hint(T + (size_t)rr * NC + pj, v[q], keep);
                        else T[(size_t)rr * NC + pj] = v[q];
                    }
                }
            }
        }
        // ---- owners: one live topic each ----
        const bool ov = owner && gt < nlive;
        const int ck = ov ? rec[2 + gt] : 0;
        const double al = ov ? p.alpha[ck] : 1.0;
        double gm = ov ? p.gam[(size_t)d * PARK_GAM + gt] : 1.0;
        double eo = ov ? exp_digamma(gm) : 0.0;
        if (owner) es[gt] = eo;
        __syncthreads();

        // ---- remaining trips on the compact tile                  (variational_bayes.py:174-190) ----
        const int M = npad / GS;
        const double* rowp = T + (size_t)(gw * LN + nl) * NC;
        const double* cntp = cs + gw * LN + nl;
        double e[CPL];
        while (true) {
#pragma unroll
            for (int i = 0; i < CPL; i += 2) {
                const double2 v = *reinterpret_cast<const double2*>(es + lcol(i));
                e[i] = v.x;
                e[i + 1] = v.y;
            }
            double s[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) s[i] = 0.0;
            int m = 0;
            if (glb) {
                for (; m + RR <= M; m += RR) longc_rows<NC, RR, true>(rowp + (size_t)m * GS * NC, cntp + m * GS, e, s, keep, po);
                for (; m < M; ++m) longc_rows<NC, 1, true>(rowp + (size_t)m * GS * NC, cntp + m * GS, e, s, keep, po);
            } else {
                for (; m + RR <= M; m += RR) longc_rows<NC, RR, false>(rowp + (size_t)m * GS * NC, cntp + m * GS, e, s, keep, po);
                for (; m < M; ++m) longc_rows<NC, 1, false>(rowp + (size_t)m * GS * NC, cntp + m * GS, e, s, keep, po);
            }
            // column sums: over the row lanes by shuffles, over the warps through shared memory
#pragma unroll
            for (int o = LK; o < 32; o <<= 1) {
#pragma unroll
                for (int i = 0; i < CPL; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
            }
            if (nl == 0) {
#pragma unroll
                for (int i = 0; i < CPL; i += 2)
                    *reinterpret_cast<double2*>(spart + gw * NC + lcol(i)) = make_double2(s[i], s[i + 1]);
            }
            __syncthreads();
            if (gw == 0) {
                // owners: gamma update (:185), |d gamma| (:187), next e
                double t0 = 0.0, t1 = 0.0;
                if (owner) {
#pragma unroll
                    for (int q = 0; q < W; q += 2) {
                        t0 += spart[q * NC + gt];
                        t1 += spart[(q + 1) * NC + gt];
                    }
                }
                const double gn = fma(eo, t0 + t1, al);
                double dd = ov ? fabs(gn - gm) : 0.0;
                if (ov) gm = gn;                                          // :188
                const double en = exp_digamma(ov ? gn : 1.0);
                dd = warp_sum(dd);
                ++it;
                const bool fin = dd <= tolK || it >= p.max_iter;          // :189-190 / :174
                if (!fin) {
                    eo = ov ? en : 0.0;
                    if (owner) es[gt] = eo;
                }
                if (gt == 0) ctl[1] = fin ? 1 : 0;
            }
            __syncthreads();
            if (ctl[1]) break;
        }

        // ---- final pass: phi from the LAST e (:207), the row weights for k_dead_phi, sum_n c_n logsumexp_n ----
        int cols[CPL];
        double ej[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int j = lcol(i);
            cols[i] = (j < nlive) ? rec[2 + j] : -1;
            ej[i] = (cols[i] >= 0) ? e[i] - p.e_dead[cols[i]] : 0.0;
        }
        double lacc = 0.0, ws = 0.0;
        for (int m = 0; m < M; ++m) {
            const int r = m * GS + gw * LN + nl;
            double b0[CPL];
            const double* r0 = rowp + (size_t)m * GS * NC;
#pragma unroll
            for (int i = 0; i < CPL; i += 2) {
                const double2 v0 = glb ? ldg_hint_v2(r0 + po[i >> 1], once)
                                       : *reinterpret_cast<const double2*>(r0 + po[i >> 1]);   // last use
                b0[i] = v0.x; b0[i + 1] = v0.y;
            }
            const double c0 = cntp[m * GS];
            const int id = is[r];
            double p0 = 0.0, q0 = 0.0;
#pragma unroll
            for (int i = 0; i < CPL; i += 2) {
                p0 = fma(b0[i], e[i], p0);
                q0 = fma(b0[i + 1], e[i + 1], q0);
            }
            p0 += q0;
#pragma unroll
            for (int o = 1; o < LK; o <<= 1) p0 += __shfl_xor_sync(0xffffffffu, p0, o);
            const bool ok = c0 > 0.0;
            const double w0 = c0 * rcp_nr(ok ? p0 : 1.0);
            if (ok) {
#pragma unroll
                for (int i = 0; i < CPL; ++i)
                    if (cols[i] >= 0) atomicAdd(p.phi_ss + (size_t)id * KP + cols[i], w0 * b0[i] * ej[i]);
                if (kl == 0) {
                    atomicAdd(p.wsum + id, w0);
                    lacc = fma(c0, p.mw[id] + log(p0), lacc);
                    ws += w0;
                }
            }
        }
        // owners: gamma of the live topics, their ELBO terms
        double t1 = lacc, sgd = 0.0;
        if (ov) {
            const double dk = gm - al;
            p.gamma[(size_t)d * K + ck] = gm;                                 // :212 / :216
            if (dk != 0.0) {
                t1 += lgamma(gm) - lgamma(al);                                // :197 (dead topics: lgamma(alpha_k), in lg_alpha)
                if (eo > 0.0) t1 -= log(eo) * dk;                             // - sum_k psi_k sum_n c_n phi_nk
                sgd += dk;
            }
        }
        t1 = warp_sum(t1);
        sgd = warp_sum(sgd);
        ws = warp_sum(ws);
        if (lane == 0) {
            red[gw] = t1;
            red[W + gw] = sgd;
            red[2 * W + gw] = ws;
        }
        __syncthreads();
        t1 = sgd = ws = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            t1 += red[w];
            sgd += red[W + w];
            ws += red[2 * W + w];
        }
        if (gt == 0) {
            p.docterm[d] = t1 + p.lg_alpha - lgamma(p.alpha_sum + sgd);       // - lgamma(sum_k gamma_k), :197
            p.iters[d] = it;
        }
        // ---- validation of the elimination (see estep_narrow.cuh): only when the bound does not settle it ----
        if (!(ws <= p.chk_bound)) {
            for (int k = gt; k < K; k += 256) sk[k] = 0.0;
            if (gt == 0) ctl[2] = 0;
            __syncthreads();
            for (int r = gw; r < n; r += W) {
                // w_n of the last trip again: one row per warp, a lane per compact column
                double pr = 0.0;
                for (int j = lane; j < NC; j += 32) pr = fma(T[(size_t)r * NC + pcol(j, r)], es[j], pr);
                pr = warp_sum(pr);
                const double c0 = cs[r];
                const double w0 = c0 * rcp_nr(c0 > 0.0 ? pr : 1.0);
                const double* brow = p.Bt + (size_t)is[r] * KP;
                for (int k = lane; k < K; k += 32) atomicAdd(sk + k, w0 * brow[k]);
            }
            __syncthreads();
            for (int k = gt; k < K; k += 256) {
                bool live = false;
                for (int j = 0; j < nlive; ++j) live = live || rec[2 + j] == k;
                if (!live && fma(p.e_dead[k], sk[k], p.alpha[k]) != p.alpha[k]) ctl[2] = 1;
            }
            __syncthreads();
            if (gt == 0 && ctl[2] && p.revived) atomicAdd(p.revived, 1);
        }
    }
}

}  // namespace pylda
