// Per-document VB E-step kernel, register-tile generation (sm_100a).
//
// Same mathematics as estep_kernel.cuh / estep_v2.cuh (reference variational_bayes.py:159-207
// in product form).  The ncu captures of estep_v2 (profiles/r1b_*) showed the shared-memory pipe
// at 50 % and the fp64 pipe at 35 % with two warps per scheduler: every trip re-read the whole
// n_d x K tile from shared memory.  Here the tile is staged into shared memory ONCE by bulk-async
// row copies (UBLKCP), moved ONCE into registers (a lane keeps R rows x 2J columns), and the
// fixed-point trips run on registers only: per trip a lane issues 4*R*J DFMA, J LDS.128 for e and
// J STS.128 for its column partial sums.  The R rows of a lane are independent dependency chains.
// Shared memory keeps e (double buffered), the column partials and -- after the last trip --
// the c*phi rows that leave by bulk reduce-add (UBLKRED), as before.
//
// A group of W warps owns one document of at most W * LN * R rows (LN = 32/LK row lanes), row
// groups dealt round-robin to the warps;
// longer documents use estep_v2 (shared-memory tile) or the streaming kernel.
#pragma once
#include "estep_v2.cuh"
#include "estep_narrow.cuh"

namespace pylda {

// rows a lane keeps in registers: 4*J*R registers of tile
template <int J>
struct RtRows {
    static constexpr int R = (J <= 5) ? 8 : (J <= 7) ? 6 : (J <= 8) ? 5 : (J <= 13) ? 3 : 2;
};

template <int LK, int J, int W, int RR_>
struct RtCfg {
    static constexpr int LN = 32 / LK;
    static constexpr int R = RR_;
    static constexpr int KPAD = 2 * LK * J;
    static constexpr int GT = 32 * W;
    static constexpr int U = (KPAD + GT - 1) / GT;
    // butterfly levels over the row lanes before the shared-memory reduce-scatter, so that an
    // owner sums at most 32 partials (a butterfly level costs every lane 2J shuffle-adds; measured
    // 25 % of the instructions of the W = 4 kernel when one level was used at 32 partials)
    static constexpr int NB = (W * LN > 128) ? 3 : (W * LN > 64) ? 2 : (W * LN > 32) ? 1 : 0;
    static constexpr int NP = (W * LN) >> NB;
    static constexpr int CAP = W * LN * R;    // rows per document group
};

// Everything between "tile is in shared memory" and "phi rows are in shared memory" for a warp
// that holds RU row groups (RU * LN rows) of the document in registers.
template <int LK, int J, int W, int RMAX, int RU>
__device__ __forceinline__ int rt_trips(const EParams& p, double* es2, double* spart, double* red, const double* cnt,
                                        const double* mwr, double* tile, int n, int g, int gt, int gw, int lane,
                                        bool warp_owns, const double* als, double* gams, double& lacc_out, int d,
                                        int& park_nl) {
    using C = RtCfg<LK, J, W, RMAX>;
    constexpr int LN = C::LN, R = C::R, KPAD = C::KPAD, GT = C::GT, U = C::U, NB = C::NB, NP = C::NP;
    constexpr int RA = RU > 0 ? RU : 1;
    const int kl = lane % LK, nl = lane / LK;
    const int K = p.K, ST = p.ST, KP2 = p.KP >> 1;
    // Row groups (LN rows each) are dealt round-robin to the W warps, so that every warp of the
    // group holds about the same number of rows: lane's rows are (i*W + gw)*LN + nl, i < RU.
    const int rbase = gw * LN + nl;
    constexpr int RSTEP = W * LN;

    // ---- tile -> registers (once per document) ----------------------------------------------
    double b[RA][2 * J];
#pragma unroll
    for (int i = 0; i < RU; ++i) {
        const int r = rbase + i * RSTEP;
        const double* rowp = tile + (size_t)(r < n ? r : 0) * ST + 2 * kl;   // rows >= n: a valid row with weight 0
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const double2 v = *reinterpret_cast<const double2*>(rowp + 2 * LK * j);
            b[i][2 * j] = v.x;
            b[i][2 * j + 1] = v.y;
        }
    }

    // Dead-topic elimination.  Once gamma_k == alpha_k bit for bit, topic k adds less
    // than half an ulp to every norm and to its own gamma (that IS the condition), e_k stops changing and
    // the topic never comes back (measured: 78 % of the topics by trip 10, 90 % by trip 20 at the headline
    // config, no revival in any corpus tried; a revival is detected in the final pass and counted).  When
    // at most 32 topics are still alive the trips continue on a compact 32-column copy of the tile:
    // 4*R*CJ instead of 4*R*J DFMA per lane and one exp(psi) pass instead of U.
    constexpr bool COMPACT = (LK >= 4) && (LK <= 16) && (KPAD > 32) && (W * LN * 32 + 48 <= NP * KPAD);
    constexpr int CJ = COMPACT ? 16 / LK : 1;           // compact topic pairs per lane (32 columns)
    bool go_compact = false;
    int nlive = 0;

    double w[RA], part[RA];
    int it = 0;
    const double tolK = p.tol * (double)K;
    while (true) {
        const double* es = es2 + (it & 1) * KPAD;
        // norm_n = B[n,:] . e : 2 chains per row, RU rows
        double a0[RA], a1[RA];
#pragma unroll
        for (int i = 0; i < RU; ++i) a0[i] = a1[i] = 0.0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const double2 ev = *reinterpret_cast<const double2*>(es + 2 * (kl + LK * j));
#pragma unroll
            for (int i = 0; i < RU; ++i) {
                a0[i] = fma(b[i][2 * j], ev.x, a0[i]);
                a1[i] = fma(b[i][2 * j + 1], ev.y, a1[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < RU; ++i) part[i] = a0[i] + a1[i];
#pragma unroll
        for (int o = 1; o < LK; o <<= 1) {
#pragma unroll
            for (int i = 0; i < RU; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], o);
        }
#pragma unroll
        for (int i = 0; i < RU; ++i) w[i] = cnt[rbase + i * RSTEP] * rcp_nr(part[i]);   // counts: staged as 0 for rows >= n
        // column partial sums of this lane -> (butterfly over NB row-lane bits) -> shared memory.
        // JB topic pairs at a time, rows in the outer loop: 2*JB independent accumulation chains.
        constexpr int JB = 1;
#pragma unroll
        for (int j0 = 0; j0 < J; j0 += JB) {
            double s0[JB], s1[JB];
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) s0[jj] = s1[jj] = 0.0;
#pragma unroll
            for (int i = 0; i < RU; ++i) {
#pragma unroll
                for (int jj = 0; jj < JB; ++jj) {
                    if (j0 + jj < J) {
                        s0[jj] = fma(w[i], b[i][2 * (j0 + jj)], s0[jj]);
                        s1[jj] = fma(w[i], b[i][2 * (j0 + jj) + 1], s1[jj]);
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) {
                if (j0 + jj < J) {
                    const int j = j0 + jj;
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        s0[jj] += __shfl_xor_sync(0xffffffffu, s0[jj], 16 >> q);
                        s1[jj] += __shfl_xor_sync(0xffffffffu, s1[jj], 16 >> q);
                    }
                    if (NB == 0 || nl < (LN >> NB))
                        *reinterpret_cast<double2*>(spart + (gw * (LN >> NB) + nl) * KPAD + 2 * (kl + LK * j)) =
                            make_double2(s0[jj], s1[jj]);
                }
            }
        }
        gsync<W>(g);
        // owners: gamma update (:185), |d gamma| (:187), speculative e for the next trip
        double gn[U], en[U];
        unsigned alive[U];
        double dsum = 0.0;
        nlive = 0;
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
            // alpha_k and gamma_k live in shared memory, not registers: the tile already takes 4*J*R
            // registers per lane and the exp(psi) evaluations below need the rest to overlap
            const double al = (k < K) ? als[k] : 1.0;
            const double ek = (k < K) ? es[k] : 0.0;                     // e_k of THIS trip (current buffer)
            gn[u] = fma(ek, (ss0 + ss1) + (ss2 + ss3), al);
            if (k < K) {
                dsum += fabs(gn[u] - gams[k]);
                gams[k] = gn[u];                                         // :188
            }
            if (COMPACT) {
                alive[u] = __ballot_sync(0xffffffffu, k < K && gn[u] != al);
                nlive += __popc(alive[u]);
            } else {
                alive[u] = 0;
            }
        }
        if (warp_owns) {
#pragma unroll
            for (int u = 0; u < U; ++u) en[u] = exp_digamma(gn[u]);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) en[u] = 0.0;
        }
        ++it;
        {
            double* esn = es2 + (it & 1) * KPAD;                         // the buffer the NEXT trip reads
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                if (k < K) esn[k] = en[u];
            }
        }
        dsum = warp_sum(dsum);
        int slot0 = 0;                                                   // first compact slot of this warp
        if (W > 1) {
            if (lane == 0) {
                red[gw] = dsum;
                red[3 * W + gw] = (double)nlive;      // (red[W .. 3W) belongs to the ELBO exchange at the end)
            }
            gsync<W>(g);
            dsum = 0.0;
            nlive = 0;
#pragma unroll
            for (int x = 0; x < W; ++x) {
                dsum += red[x];
                if (x < gw) slot0 += (int)red[3 * W + x];
                nlive += (int)red[3 * W + x];
            }
        } else {
            __syncwarp();
        }
        if (dsum <= tolK || it >= p.max_iter) break;                     // :189-190 / :174
        if (COMPACT && p.compact && nlive <= 32) {
            // ---- switch: both e buffers get the current e of every topic (dead ones keep it for good),
            // the live topics are numbered 0 .. nlive-1 in topic order
            double* eso = es2 + ((it + 1) & 1) * KPAD;
            // (the column partials in spart have been consumed: every warp is past the second barrier)
            int* livecol = reinterpret_cast<int*>(spart + W * LN * 32);
            double* es_c = spart + W * LN * 32 + 16;
            int base_slot = slot0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                if (k < K) eso[k] = en[u];
                if ((alive[u] >> lane) & 1u) {
                    const int slot = base_slot + __popc(alive[u] & ((1u << lane) - 1u));
                    livecol[slot] = k;
                    es_c[slot] = en[u];
                }
                base_slot += __popc(alive[u]);
            }
            if (gw == 0 && lane >= nlive) {
                livecol[lane] = 0;
                es_c[lane] = 0.0;
            }
            gsync<W>(g);
            go_compact = true;
            break;
        }
    }

    if (COMPACT && go_compact) {
        int* livecol = reinterpret_cast<int*>(spart + W * LN * 32);
        double* es_c = spart + W * LN * 32 + 16;
        double* spart_c = spart;                                          // [W * LN][32]
        // the lane's compact tile: RU rows x 2*CJ live columns, gathered from the staged tile
        double bc[RA][2 * CJ];
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj) {
            const int c0 = livecol[2 * (kl + LK * jj)], c1 = livecol[2 * (kl + LK * jj) + 1];
#pragma unroll
            for (int i = 0; i < RU; ++i) {
                const int r = rbase + i * RSTEP;
                const double* rowp = tile + (size_t)(r < n ? r : 0) * ST;
                bc[i][2 * jj] = rowp[c0];
                bc[i][2 * jj + 1] = rowp[c1];
            }
        }
        // lane t of the group's first warp owns live slot t
        const bool valid = (gw == 0) && lane < nlive;
        const int kt = livecol[lane];
        const double alt = als[kt];
        double gcur = gams[kt];
        double et = es_c[lane];
        while (true) {
            double a0[RA], a1[RA];
#pragma unroll
            for (int i = 0; i < RU; ++i) a0[i] = a1[i] = 0.0;
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                const double2 ev = *reinterpret_cast<const double2*>(es_c + 2 * (kl + LK * jj));
#pragma unroll
                for (int i = 0; i < RU; ++i) {
                    a0[i] = fma(bc[i][2 * jj], ev.x, a0[i]);
                    a1[i] = fma(bc[i][2 * jj + 1], ev.y, a1[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < RU; ++i) part[i] = a0[i] + a1[i];
#pragma unroll
            for (int o = 1; o < LK; o <<= 1) {
#pragma unroll
                for (int i = 0; i < RU; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], o);
            }
#pragma unroll
            for (int i = 0; i < RU; ++i) w[i] = cnt[rbase + i * RSTEP] * rcp_nr(part[i]);
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int i = 0; i < RU; ++i) {
                    s0 = fma(w[i], bc[i][2 * jj], s0);
                    s1 = fma(w[i], bc[i][2 * jj + 1], s1);
                }
                *reinterpret_cast<double2*>(spart_c + (gw * LN + nl) * 32 + 2 * (kl + LK * jj)) = make_double2(s0, s1);
            }
            gsync<W>(g);
            ++it;
            bool done = false;
            if (gw == 0) {
                double ss0 = 0.0, ss1 = 0.0, ss2 = 0.0, ss3 = 0.0;
#pragma unroll
                for (int q = 0; q < W * LN; q += 4) {
                    ss0 += spart_c[q * 32 + lane];
                    if (q + 1 < W * LN) ss1 += spart_c[(q + 1) * 32 + lane];
                    if (q + 2 < W * LN) ss2 += spart_c[(q + 2) * 32 + lane];
                    if (q + 3 < W * LN) ss3 += spart_c[(q + 3) * 32 + lane];
                }
                const double gnc = fma(et, (ss0 + ss1) + (ss2 + ss3), alt);   // :185
                double dsum = valid ? fabs(gnc - gcur) : 0.0;             // :187 (dead topics contribute exactly 0)
                if (valid) gcur = gnc;                                    // :188
                const double enc = exp_digamma(valid ? gnc : 1.0);
                dsum = warp_sum(dsum);
                done = dsum <= tolK || it >= p.max_iter;                  // :189-190 / :174
                if (!done) {
                    et = enc;
                    es_c[lane] = valid ? enc : 0.0;
                    if (p.park_nc > 0 && n <= 192) {
                        // few enough topics alive: park the document for the narrow stages (estep_narrow.cuh)
                        const unsigned lvb = __ballot_sync(0xffffffffu, valid && gnc != alt);
                        const int nl2 = __popc(lvb);
                        if (nl2 >= 1 && nl2 <= p.park_nc) {
                            int* rec = p.park_rec + (size_t)d * PARK_REC;
                            if ((lvb >> lane) & 1u) {
                                const int rank = __popc(lvb & ((1u << lane) - 1u));
                                rec[2 + rank] = kt;
                                p.park_gam[(size_t)d * PARK_GAM + rank] = gnc;
                            }
                            if (lane == 0) {
                                rec[0] = it;
                                rec[1] = nl2;
                            }
                            park_nl = nl2;
                            done = true;
                        }
                    }
                }
                if (W > 1 && lane == 0) {
                    red[0] = done ? 1.0 : 0.0;
                    red[1] = (double)park_nl;
                }
            }
            gsync<W>(g);
            if (W > 1) {
                done = red[0] != 0.0;
                park_nl = (int)red[1];
            }
            if (done) break;
        }
        // back to the full-width arrays: gamma, and the e of the LAST trip into the buffer the final pass reads
        if (valid) {
            gams[kt] = gcur;
            es2[((it - 1) & 1) * KPAD + kt] = et;
        }
        gsync<W>(g);
        if (park_nl > 0) {                                                // parked: the narrow stages finish it
            lacc_out = 0.0;
            return it;
        }
    }

    // ---- phi from the LAST e (buffer (it-1)&1; w[] and part[] are those of the last trip) -------
    const double* es = es2 + ((it - 1) & 1) * KPAD;
    double lacc = 0.0;
#pragma unroll
    for (int i = 0; i < RU; ++i) {
        const int r = rbase + i * RSTEP;
        if (r < n && kl == 0) lacc = fma(cnt[r], mwr[r] + log(part[i]), lacc);   // sum_n c_n logsumexp_n
    }
    bool came_back = false;
    if (COMPACT && go_compact) {
        // Validation of the elimination: full-width column sums with the weights of the last trip; every
        // eliminated topic must still satisfy alpha_k + e_k s_k == alpha_k.  (A revival has never been
        // observed; it is counted and reported, see pylda_stats.revived_docs.)
#pragma unroll
        for (int j = 0; j < J; ++j) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int i = 0; i < RU; ++i) {
                const int r = rbase + i * RSTEP;
                const double2 bv = *reinterpret_cast<const double2*>(tile + (size_t)(r < n ? r : 0) * ST + 2 * (kl + LK * j));
                s0 = fma(w[i], bv.x, s0);
                s1 = fma(w[i], bv.y, s1);
            }
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, 16 >> q);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 16 >> q);
            }
            if (NB == 0 || nl < (LN >> NB))
                *reinterpret_cast<double2*>(spart + (gw * (LN >> NB) + nl) * KPAD + 2 * (kl + LK * j)) = make_double2(s0, s1);
        }
        gsync<W>(g);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = gt + GT * u;
            if (k < K && gams[k] == als[k]) {
                double ss = 0.0;
#pragma unroll
                for (int q = 0; q < NP; ++q) ss += spart[q * KPAD + k];
                came_back |= fma(es[k], ss, als[k]) != als[k];
            }
        }
        if (came_back && p.revived) atomicAdd(p.revived, 1);
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const double2 ev = *reinterpret_cast<const double2*>(es + 2 * (kl + LK * j));
        if (kl + LK * j < KP2) {
#pragma unroll
            for (int i = 0; i < RU; ++i) {
                const int r = rbase + i * RSTEP;
                if (r < n) {
                    double2* cell = reinterpret_cast<double2*>(tile + (size_t)r * ST + 2 * (kl + LK * j));
                    const double2 bv = COMPACT ? *cell : make_double2(b[i][2 * j], b[i][2 * j + 1]);
                    *cell = make_double2(w[i] * bv.x * ev.x, w[i] * bv.y * ev.y);   // c_n phi_nk (:207)
                }
            }
        }
    }
    lacc_out = lacc;
    return it;
}

// RMAX = rows a lane keeps in registers, NWARPS = warps per CTA (register budget 65536 / (32 NWARPS)):
// the default (RtRows<J>::R, 8 warps, 255 registers) holds the most rows per warp; smaller RMAX with 12 or
// 16 warps per CTA trades rows per warp for warps per scheduler (a lone warp needs ~2700 cycles per trip
// for ~850 cycles of fp64 pipe: the kernels are latency bound, so more resident warps is more throughput).
template <int LK, int J, int W, int RMAX, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) estep_rt(const EParams p) {
    using C = RtCfg<LK, J, W, RMAX>;
    constexpr int LN = C::LN, R = C::R, KPAD = C::KPAD, GT = C::GT, U = C::U, CAP = C::CAP;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int g = tid / GT;
    const int gt = tid - g * GT;
    const int gw = gt >> 5;
    const int lane = tid & 31;
    const int K = p.K, KP = p.KP, ST = p.ST;

    unsigned char* gs = smem_raw + (size_t)g * p.group_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(gs);
    int* cur = reinterpret_cast<int*>(gs + 8);
    double* es2 = reinterpret_cast<double*>(gs + 16);
    double* spart = reinterpret_cast<double*>(gs + p.off_spart);
    double* red = reinterpret_cast<double*>(gs + p.off_red);
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

    // alpha: one copy per CTA behind the groups; gamma_k of the group's document: in the slack behind
    // its tile (finite values, which is all the over-read of the last tile row needs)
    double* als = reinterpret_cast<double*>(smem_raw + p.off_groups);
    double* gams = tile + (size_t)CAP * ST;
    for (int k = tid; k < KPAD; k += blockDim.x) als[k] = (k < K) ? p.alpha[k] : 1.0;
    __syncthreads();

    const bool warp_owns = gw * 32 < K;
    uint32_t parity = 0;
    int nxt = 0;
    if (gt == 0) nxt = atomicAdd(p.counter, 1);

    while (true) {
        bulk_wait_read0();
        int idx;
        if (W == 1) {
            idx = __shfl_sync(0xffffffffu, nxt, 0);
        } else {
            if (gt == 0) *cur = nxt;
            gsync<W>(g);
            idx = *cur;
        }
        if (idx >= p.ndocs) break;
        if (gt == 0) nxt = atomicAdd(p.counter, 1);
        const int d = p.order[idx];
        const long long base = p.row_ptr[d];
        const int n = (int)(p.row_ptr[d + 1] - base);     // host guarantees n <= CAP for this class

        if (gt == 0) mbar_expect_tx(mbar, (uint32_t)n * (uint32_t)KP * 8u);
        int csum = 0;
        for (int r = gt; r < CAP; r += GT) {
            int c = 0;
            if (r < n) {
                const int id = p.ids[base + r];
                c = p.cts[base + r];
                rid[r] = id;
                mwr[r] = p.mw[id];
                bulk_g2s(tile + (size_t)r * ST, p.Bt + (size_t)id * KP, (uint32_t)KP * 8u, mbar);
            }
            cnt[r] = (double)c;
            csum += c;
        }
        csum = __reduce_add_sync(0xffffffffu, csum);
        double Nd = (double)csum;
        if (W > 1) {
            if (lane == 0) red[gw] = Nd;
            gsync<W>(g);
            Nd = 0.0;
#pragma unroll
            for (int x = 0; x < W; ++x) Nd += red[x];
        }
        const double g0 = Nd / (double)K;                  // gamma0 = alpha + N_d / K   (:165)
        if (warp_owns) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                const double g00 = ((k < K) ? als[k] : 1.0) + g0;
                if (k < K) gams[k] = g00;
                const double e0 = exp_digamma(g00);
                if (k < K) es2[k] = e0;
            }
        }
        mbar_wait(mbar, parity);
        parity ^= 1u;
        gsync<W>(g);

        // row groups of this warp: gw, gw + W, ... below ceil(n / LN)
        const int NG = (n + LN - 1) / LN;
        int RU = (NG - gw + W - 1) / W;
        RU = RU < 0 ? 0 : (RU > R ? R : RU);
        double lacc = 0.0;
        int it = 0, park_nl = 0;
#define PYLDA_RT_CASE(X)                                                                                     \
    case X:                                                                                                  \
        if constexpr (X <= R)                                                                                \
            it = rt_trips<LK, J, W, RMAX, X>(p, es2, spart, red, cnt, mwr, tile, n, g, gt, gw, lane, warp_owns, als, \
                                       gams, lacc, d, park_nl);                                              \
        break;
        switch (RU) {
            PYLDA_RT_CASE(0) PYLDA_RT_CASE(1) PYLDA_RT_CASE(2) PYLDA_RT_CASE(3) PYLDA_RT_CASE(4)
            PYLDA_RT_CASE(5) PYLDA_RT_CASE(6) PYLDA_RT_CASE(7) PYLDA_RT_CASE(8)
        }
#undef PYLDA_RT_CASE
        fence_async_smem();   // generic-proxy writes of phi -> visible to the bulk-async engine
        if (park_nl > 0) {
            // parked: gamma so far (final for every dead topic) and the document joins its narrow-stage list
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = gt + GT * u;
                if (k < K) p.gamma[(size_t)d * K + k] = gams[k];
            }
            if (gt == 0) {
                const int li = park_list_index(park_nl, n);
                const int slot = atomicAdd(p.park_counts + li, 1);
                p.park_lists[(size_t)li * p.park_cap + slot] = d;
            }
            gsync<W>(g);
            continue;
        }

        // ---- per-document ELBO pieces and gamma write-back ------------------------------------
        double t1 = lacc, sg = 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = gt + GT * u;
            if (k < K) {
                const double gk = gams[k];
                const double ek = es2[((it - 1) & 1) * KPAD + k];           // e of the last trip
                const double dk = gk - als[k];
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
                for (int x = 0; x < W; ++x) {
                    t1 += red[W + 2 * x];
                    sg += red[W + 2 * x + 1];
                }
            }
            p.docterm[d] = t1 - lgamma(sg);                                  // - lgamma(sum_k gamma_k), :197
            p.iters[d] = it;
        }
    }
}

}  // namespace pylda
