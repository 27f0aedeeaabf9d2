// Per-document VB E-step kernel, "narrow" stages (sm_100a): documents whose live topics fit 16 / 8 columns.
//
// Same mathematics as the other generations (reference variational_bayes.py:174-207 in product form).  Why it
// exists: at the cold state of the headline config a short document spends 30-40 of its ~44 trips with
// fewer than 8 topics alive (gamma_k != alpha_k; scripts/sim_live.py), yet the register-tile kernel keeps a
// whole warp (or W warps) on it and pays ~250 warp instructions per trip for ~10 useful DFMA
// (profiles/r1f_estep_rt_block_view.txt: the 32-column compact trip block is 40 % of the one-warp kernel).
// estep_rt therefore PARKS a document once at most 16 (n <= 96) or 8 topics are alive: it writes gamma
// (final for every dead topic), the trip count and the live columns, and moves on.  The kernels here pick the
// parked documents up: G = 4..32 lanes per document (32/G documents per warp), a lane keeps RPL rows x NC live
// columns of the document's tile in registers (48 doubles: 6 x 8 or 3 x 16), gathered directly from the
// (V, KP) table.  A trip is RPL*NC DFMA for the norms (no cross-lane reduction: a lane owns whole rows),
// RPL reciprocals, RPL*NC DFMA for the column sums, one reduce-scatter of NC values through shared memory,
// NC/G exp(psi) per owner lane -- about 35 (G = 4) to 60 (G = 8) warp instructions per document-trip.
// The 16-column stage hands a document over to the 8-column stage the same way.
//
// phi: for a parked document every topic outside the live columns has e_k = exp(psi(alpha_k)) =: ed_k for good, so
//   c_n phi_nk = w_n B[n,k] e_k = w_n B[n,k] ed_k + [k live] w_n B[n,k] (e_k - ed_k).
// The second term is scattered from registers (NC columns per row, red.global.add.f64); the first is summed
// over all parked documents per WORD: the kernels add w_n to wsum[word] (one red per row instead of K) and
// k_dead_phi adds ed_k B[w,k] wsum[w] to the statistics in one pass over the (V, KP) table.  Every entry of
// phi_ss receives exactly what the reference adds (variational_bayes.py:207), re-associated.
#pragma once
#include "estep_kernel.cuh"

namespace pylda {

constexpr int PARK_REC = 36;    // ints per parked document: [0] trips done, [1] live topics, [2..34) their columns
constexpr int PARK_GAM = 32;    // doubles per parked document: gamma of the live topics (same order)
constexpr int PARK_LISTS = 10;  // 16-column stage: G = 8, 16, 32 (n <= 24, 48, 96) [0..2] and G = 32 with 6 rows per lane
                                // (n <= 192) [7]; 8-column stage: G = 4, 8, 16, 32 (n <= 24, 48, 96, 192) [3..6];
                                // 32-column stage: G = 32 (n <= 96) [8], fed by the kernels without a compact stage;
                                // long documents (n > 192, at most 32 topics alive) [9]: estep_longc.cuh
constexpr int PARK_NARROW_LISTS = 9;   // lists served by estep_narrow
constexpr int PARK_LONG_MIN_TRIPS = 10; // a long document is handed over only while at least this many trips are left
constexpr int PARK_CTRS = 32;   // ints: list lengths [0, 16) and queue heads [16, 32)

// list a parked document joins: by stage (live topics) and length
__device__ __forceinline__ int park_list_index(int nlive, int n) {
    if (n > 192) return 9;
    if (nlive > 16) return 8;
    if (nlive > 8) return n <= 24 ? 0 : n <= 48 ? 1 : n <= 96 ? 2 : 7;
    return n <= 24 ? 3 : n <= 48 ? 4 : n <= 96 ? 5 : 6;
}

struct NParams {
    const long long* __restrict__ row_ptr;
    const int* __restrict__ ids;
    const int* __restrict__ cts;
    const double* __restrict__ Bt;
    const double* __restrict__ mw;
    const double* __restrict__ alpha;
    const double* __restrict__ e_dead;   // exp(psi(alpha_k)): e of an eliminated topic
    double* gamma;
    double* phi_ss;
    double* wsum;                        // (V,) sum of w_n over the parked documents' rows, per word
    double* docterm;
    int* iters;
    int K, KP, max_iter;
    double tol;
    double lg_alpha;                     // sum_k lgamma(alpha_k)
    double alpha_sum;                    // sum_k alpha_k
    const int* list;                     // documents of this class
    const int* count;
    int* head;                           // queue head
    int* rec;                            // park records (PARK_REC ints per document)
    double* gam;                         // gamma of the live topics (PARK_GAM doubles per document)
    int* lists;                          // all PARK_LISTS lists (hand-over 16 -> 8 columns)
    int* counts;
    int cap;                             // capacity of one list
    int handover;                        // hand a document over to the next narrower stage at NC / 2 live topics
    double chk_bound;                    // sum_n w_n <= chk_bound proves that no eliminated topic can come back
    int* revived;                        // counter of documents in which one would have
};

// RPL rows per lane: 48 / NC (96 registers of tile), or twice that; MINB CTAs per SM (register budget 65536 / (128 MINB))
template <int NC, int G, int RPL, int MINB>
__global__ void __launch_bounds__(128, MINB) estep_narrow(const NParams p) {
    constexpr int NG = 32 / G;                        // documents per warp
    constexpr int SPL = (NC >= G) ? NC / G : 1;       // topic slots per owner lane
    constexpr int NOWN = (NC >= G) ? G : NC;          // owner lanes per document
    constexpr int PST = NC + 2;                       // padded stride of a lane's column partials (doubles)
    constexpr int WSM = 32 * PST + NG * NC + NG * 16; // doubles per warp: partials, e, live-column bit masks
    extern __shared__ __align__(16) double nsm[];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane / G, gl = lane % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (g * G));
    double* sp = nsm + (size_t)wid * WSM;
    double* se = sp + 32 * PST + g * NC;
    unsigned* lm = reinterpret_cast<unsigned*>(sp + 32 * PST + NG * NC) + g * 32;
    const int K = p.K, KP = p.KP;
    const double tolK = p.tol * (double)K;
    const int ndocs = *p.count;
    const bool owner = gl < NOWN;

    while (true) {
        int b0 = 0;
        if (lane == 0) b0 = atomicAdd(p.head, NG);
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        if (b0 >= ndocs) break;
        const int idx = b0 + g;
        const bool active = idx < ndocs;
        int d = 0, n = 0, it = 0, nlive = 0;
        long long base = 0;
        const int* rec = p.rec;
        if (active) {
            d = p.list[idx];
            base = p.row_ptr[d];
            n = (int)(p.row_ptr[d + 1] - base);
            rec = p.rec + (size_t)d * PARK_REC;
            it = rec[0];
            nlive = rec[1];
        }
        // ---- the lane's rows: RPL rows x NC live columns, gathered from the (V, KP) table ----
        double bt[RPL][NC], c[RPL];
        int id[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int r = gl + G * i;
            const bool ok = active && r < n;
            id[i] = ok ? p.ids[base + r] : 0;
            c[i] = ok ? (double)p.cts[base + r] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int col = (active && j < nlive) ? rec[2 + j] : 0;
#pragma unroll
            for (int i = 0; i < RPL; ++i) bt[i][j] = p.Bt[(size_t)id[i] * KP + col];
        }
        // ---- owner lanes: SPL topic slots each ----
        double al[SPL], gm[SPL], eo[SPL];
        int ck[SPL];
        bool ov[SPL];
#pragma unroll
        for (int u = 0; u < SPL; ++u) {
            const int slot = gl * SPL + u;
            ov[u] = owner && active && slot < nlive;
            ck[u] = ov[u] ? rec[2 + slot] : 0;
            al[u] = ov[u] ? p.alpha[ck[u]] : 1.0;
            gm[u] = ov[u] ? p.gam[(size_t)d * PARK_GAM + slot] : 1.0;
            eo[u] = ov[u] ? exp_digamma(gm[u]) : 0.0;
            if (owner) se[slot] = eo[u];
        }
        __syncwarp();

        double w[RPL], part[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i) w[i] = 0.0, part[i] = 1.0;
        // The trip loop is WARP-uniform: it runs until every document of the warp has stopped.  A stopped document
        // no longer updates gamma, e or its trip counter, so its extra trips recompute the same norms and weights
        // (idempotent) -- no per-group control flow, full-mask __syncwarp / shuffles.  (Per-group masks made ptxas
        // emit MATCH / BRA.DIV sequences and the groups ended up running the loop body separately: 2.3 passes per
        // trip in profiles/r2c_narrow_blocks.txt.)
        bool fin = !active, parked = false;
        while (true) {
            // norm_n = B[n, live] . e   (a lane owns whole rows: no cross-lane reduction)
            {
                double a0[RPL], a1[RPL];
#pragma unroll
                for (int i = 0; i < RPL; ++i) a0[i] = a1[i] = 0.0;
#pragma unroll
                for (int j = 0; j < NC; j += 2) {
                    const double2 ev = *reinterpret_cast<const double2*>(se + j);
#pragma unroll
                    for (int i = 0; i < RPL; ++i) {
                        a0[i] = fma(bt[i][j], ev.x, a0[i]);
                        a1[i] = fma(bt[i][j + 1], ev.y, a1[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < RPL; ++i) part[i] = a0[i] + a1[i];
            }
#pragma unroll
            for (int i = 0; i < RPL; ++i) w[i] = c[i] * rcp_nr(c[i] > 0.0 ? part[i] : 1.0);   // branch-free: the reciprocals interleave
            // column partial sums of this lane -> shared memory
#pragma unroll
            for (int j = 0; j < NC; j += 2) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int i = 0; i < RPL; ++i) {
                    s0 = fma(w[i], bt[i][j], s0);
                    s1 = fma(w[i], bt[i][j + 1], s1);
                }
                *reinterpret_cast<double2*>(sp + lane * PST + j) = make_double2(s0, s1);
            }
            __syncwarp();
            // owners: gamma update (:185), |d gamma| (:187), next e
            double dd = 0.0, en[SPL];
            bool lv[SPL];
#pragma unroll
            for (int u = 0; u < SPL; ++u) {
                const int slot = gl * SPL + u;
                double t0 = 0.0, t1 = 0.0;
                if (owner) {
#pragma unroll
                    for (int q = 0; q < G; q += 2) {
                        t0 += sp[(g * G + q) * PST + slot];
                        t1 += sp[(g * G + q + 1) * PST + slot];
                    }
                }
                const double gn = fma(eo[u], t0 + t1, al[u]);
                if (ov[u] && !fin) {
                    dd += fabs(gn - gm[u]);
                    gm[u] = gn;                                           // :188
                }
                lv[u] = ov[u] && gn != al[u];
                en[u] = exp_digamma(ov[u] ? gn : 1.0);
            }
#pragma unroll
            for (int o = G >> 1; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
            const bool running = !fin;
            if (running) {
                ++it;
                fin = dd <= tolK || it >= p.max_iter;                     // :189-190 / :174
            }
            if (NC > 8) {
                // hand-over to the next narrower stage once at most NC / 2 topics are alive
                unsigned bal[SPL];
                int nl = 0;
#pragma unroll
                for (int u = 0; u < SPL; ++u) {
                    bal[u] = __ballot_sync(0xffffffffu, lv[u]) & gmask;
                    nl += __popc(bal[u]);
                }
                if (p.handover && running && !fin && nl <= NC / 2) {
                    int* wrec = p.rec + (size_t)d * PARK_REC;
                    const unsigned below = gmask & ((1u << lane) - 1u);
                    int rank = 0;
#pragma unroll
                    for (int u = 0; u < SPL; ++u) rank += __popc(bal[u] & below);
#pragma unroll
                    for (int u = 0; u < SPL; ++u) {
                        if (ov[u]) p.gamma[(size_t)d * K + ck[u]] = gm[u];   // final for the topics that just died
                        if (lv[u]) {
                            wrec[2 + rank] = ck[u];
                            p.gam[(size_t)d * PARK_GAM + rank] = gm[u];
                            ++rank;
                        }
                    }
                    if (gl == 0) {
                        wrec[0] = it;
                        wrec[1] = nl;
                        const int li = park_list_index(nl, n);
                        const int slot = atomicAdd(p.counts + li, 1);
                        p.lists[(size_t)li * p.cap + slot] = d;
                    }
                    parked = true;
                    fin = true;
                }
            }
            if (!fin) {
#pragma unroll
                for (int u = 0; u < SPL; ++u) {
                    eo[u] = ov[u] ? en[u] : 0.0;
                    if (owner) se[gl * SPL + u] = eo[u];
                }
            }
            __syncwarp();
            if (!__any_sync(0xffffffffu, !fin)) break;
        }

        // ---- epilogue (all documents of the warp together; w / part / se are those of the last trip) ----
        const bool fini = active && !parked;
        double t1 = 0.0, sgd = 0.0;
        if (fini) {
#pragma unroll
            for (int i = 0; i < RPL; ++i)
                if (c[i] > 0.0) t1 = fma(c[i], p.mw[id[i]] + log(part[i]), t1);   // sum_n c_n logsumexp_n
#pragma unroll
            for (int u = 0; u < SPL; ++u) {
                if (ov[u]) {
                    const double dk = gm[u] - al[u];
                    p.gamma[(size_t)d * K + ck[u]] = gm[u];                   // :212 / :216
                    if (dk != 0.0) {
                        t1 += lgamma(gm[u]) - lgamma(al[u]);                  // :197 (dead topics: lgamma(alpha_k), in lg_alpha)
                        if (eo[u] > 0.0) t1 -= log(eo[u]) * dk;               // - sum_k psi_k sum_n c_n phi_nk
                        sgd += dk;
                    }
                }
            }
        }
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) {
            t1 += __shfl_xor_sync(0xffffffffu, t1, o);
            sgd += __shfl_xor_sync(0xffffffffu, sgd, o);
        }
        if (fini && gl == 0) {
            p.docterm[d] = t1 + p.lg_alpha - lgamma(p.alpha_sum + sgd);       // - lgamma(sum_k gamma_k), :197
            p.iters[d] = it;
        }
        if (fini) {
            // c_n phi_nk (:207): live columns from registers (minus the dead-topic value that k_dead_phi adds for
            // every column), and the row weights for k_dead_phi
            // (all loads first: behind the reduce-adds they would each wait a full L2 round trip)
            int cols[NC];
            double ej[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) cols[j] = (j < nlive) ? rec[2 + j] : 0;
#pragma unroll
            for (int j = 0; j < NC; ++j) ej[j] = se[j] - p.e_dead[cols[j]];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                if (j < nlive) {
#pragma unroll
                    for (int i = 0; i < RPL; ++i)
                        if (c[i] > 0.0) atomicAdd(p.phi_ss + (size_t)id[i] * KP + cols[j], w[i] * bt[i][j] * ej[j]);
                }
            }
#pragma unroll
            for (int i = 0; i < RPL; ++i)
                if (c[i] > 0.0) atomicAdd(p.wsum + id[i], w[i]);
        }
        // Validation of the elimination: a dead topic stays dead while alpha_k + e_k s_k == alpha_k, s_k =
        // sum_n w_n B[n,k] <= sum_n w_n (B <= 1).  The bound is checked first; only when it fails the column
        // sums of the eliminated topics are formed.
        double ws = 0.0;
#pragma unroll
        for (int i = 0; i < RPL; ++i) ws += w[i];
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
        const bool full_pass = fini && !(ws <= p.chk_bound);
        if (__any_sync(0xffffffffu, full_pass)) {
            for (int x = gl; x < 32; x += G) lm[x] = 0u;
            __syncwarp();
#pragma unroll
            for (int u = 0; u < SPL; ++u)
                if (ov[u]) atomicOr(lm + (ck[u] >> 5), 1u << (ck[u] & 31));
            __syncwarp();
            bool came_back = false;
            for (int k0 = 0; k0 < K; k0 += G) {
                const int k = k0 + gl;
                const bool dead = k < K && !((lm[k >> 5] >> (k & 31)) & 1u);
                const double ed = dead ? p.e_dead[k] : 0.0;
                double sk = 0.0;
#pragma unroll
                for (int i = 0; i < RPL; ++i) {
                    for (int q = 0; q < G; ++q) {
                        const double wq = __shfl_sync(0xffffffffu, w[i], g * G + q);
                        const int idq = __shfl_sync(0xffffffffu, id[i], g * G + q);
                        if (full_pass && dead && q + G * i < n) {
                            const double bv = p.Bt[(size_t)idq * KP + k];
                            sk = fma(wq, bv, sk);
                        }
                    }
                }
                if (full_pass && dead && fma(ed, sk, p.alpha[k]) != p.alpha[k]) came_back = true;
            }
            if ((__ballot_sync(0xffffffffu, came_back) & gmask) && gl == 0 && p.revived) atomicAdd(p.revived, 1);
        }
        __syncwarp();
    }
}

}  // namespace pylda
