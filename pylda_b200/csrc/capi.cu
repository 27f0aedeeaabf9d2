// C ABI of the B200-native VB E-step (include/pylda_b200.h).  Host orchestration only:
// device memory, streams/events, class scheduling of documents, NCCL (dlopen'd) and the
// launches of the kernels in estep_kernel.cuh / prep_kernels.cuh.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/pylda_b200.h"
#include "estep_dispatch.h"
#include "estep_kernel.cuh"
#include "estep_narrow.cuh"
#include "estep_longc.cuh"
#include "estep_sweep.cuh"
#include "prep_kernels.cuh"

using namespace pylda;

namespace pylda {
int device_top_words(const double* Elt, const double* lse, int K, int V, int KP, int top, int32_t* idx_out, double* prob_out,
                     cudaStream_t s, std::string* err);
}

namespace {

std::string g_create_error;

// ---- minimal NCCL binding (resolved at pylda_comm_init; single-GPU runs never need it) ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[PYLDA_NCCL_ID_BYTES]; } ncclUniqueId;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
const int kNcclFloat64 = 8, kNcclSum = 0;

bool load_nccl(std::string* err) {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        *err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
        return false;
    }
#define SYM(field, name)                                              \
    *(void**)(&g_nccl.field) = dlsym(h, name);                        \
    if (!g_nccl.field) {                                              \
        *err = std::string("libnccl is missing symbol ") + name;      \
        return false;                                                 \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return true;
}

struct Corpus {
    long long D = 0, nnz = 0;
    long long* row_ptr = nullptr;
    int* ids = nullptr;
    int* cts = nullptr;
    int* order = nullptr;            // documents sorted by n_d, longest first
    std::vector<int> n_sorted;       // host copy of the sorted lengths
    int max_id = -1;
    // per-corpus outputs
    double* gamma = nullptr;
    size_t gamma_cap = 0;            // doubles
    double* gamma_dst = nullptr;     // where the kernels of the current E-step write gamma: `gamma`, or the
                                     // device alias of a caller's page-locked host buffer (zero-copy D2H)
    bool gamma_on_device = false;
    // early copy of gamma (estep_resident_impl): the caller's page-locked buffer and its device alias for the
    // E-step in flight; early_done = the buffer holds the complete gamma when the per-document kernels have finished
    double* early_host = nullptr;
    double* early_alias = nullptr;
    bool early_done = false;
    double* docterm = nullptr;
    int* iters = nullptr;
    bool has_results = false;
    int results_K = 0;
    // hand-over buffers of the narrow stages (estep_narrow.cuh)
    int* park_rec = nullptr;         // PARK_REC ints per document
    double* park_gam = nullptr;      // PARK_GAM doubles per document
    int* park_lists = nullptr;       // PARK_LISTS lists of D documents
};

}  // namespace

struct pylda_ctx {
    int device = 0;
    cudaDeviceProp prop;
    cudaStream_t stream = nullptr;
    double* flat_part = nullptr;             // k_build_B: per-block sums of Bt; flat_dev / flat_host: their total
    double* flat_dev = nullptr;
    double* flat_host = nullptr;             // page-locked
    cudaEvent_t ev_flat = nullptr;
    double* longc_tile = nullptr;            // scratch of estep_longc (per CTA: compact tile, counts, term ids)
    double* longc_cnt = nullptr;
    int* longc_ids = nullptr;
    size_t longc_rows = 0;
    cudaStream_t copy_stream = nullptr;      // the early D2H copy of gamma runs beside the long-document kernels
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    std::string err;
    Corpus corp[2];
    // model
    int K = 0, V = 0, KP = 0;
    double* eta = nullptr;       // (K, V)
    double* alpha = nullptr;     // (K,)
    double alpha_max = 0.0, alpha_min = 0.0, alpha_sum = 0.0, lg_alpha = 0.0, alpha_term = 0.0;
    std::vector<double> alpha_host;
    double* Elt = nullptr;       // (V, KP) E_log_eta transposed
    double* Bt = nullptr;        // (V, KP)
    double* mw = nullptr;        // (V,)
    double* phi = nullptr;       // (V, KP) statistics accumulator
    double* phi_KV = nullptr;    // (K, V) reference layout
    double* kbuf = nullptr;      // 4*K doubles: psisum, rowsum, lse, rowterm
    double* alpha_ss = nullptr;  // (K,)
    double* scal = nullptr;      // 8 doubles
    double* partial = nullptr;   // reduction scratch
    size_t partial_cap = 0;
    int* counters = nullptr;     // 32 ints: class queue heads [1..13), revived documents [14], long-document kernels [16..19)
    int* park_ctr = nullptr;     // narrow stages: list lengths [0..16) and queue heads [16..32)
    double* e_dead = nullptr;    // (K,) exp(psi(alpha_k))
    double* wsum = nullptr;      // (V,) row weights of the documents finished by the narrow stages
    float* Bt32 = nullptr;       // (V, KP) fp32 copy of Bt (precision sweep only, PYLDA_PRECISION)
    bool model_set = false;
    bool phi_KV_valid = false;
    bool have_alpha_ss = false;
    bool force_full = false;     // redo of an E-step in which an eliminated topic came back: no elimination, no hand-over
    int phi_slot = -1;           // which corpus slot / branch produced the statistics in `phi`
    bool phi_heldout = false;
    double last_scal[8] = {0};
    // comm
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
};

namespace {

int fail(pylda_ctx* c, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return 1;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t n) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    if (n == 0) n = 1;
    return cudaMalloc((void**)p, n * sizeof(T));
}

void free_corpus(Corpus& c) {
    cudaFree(c.row_ptr); cudaFree(c.ids); cudaFree(c.cts); cudaFree(c.order);
    cudaFree(c.gamma); cudaFree(c.docterm); cudaFree(c.iters);
    cudaFree(c.park_rec); cudaFree(c.park_gam); cudaFree(c.park_lists);
    c = Corpus();
}

void free_model(pylda_ctx* c) {
    cudaFree(c->eta); cudaFree(c->alpha); cudaFree(c->Elt); cudaFree(c->Bt); cudaFree(c->mw);
    cudaFree(c->phi); cudaFree(c->phi_KV); cudaFree(c->kbuf); cudaFree(c->alpha_ss); cudaFree(c->e_dead); cudaFree(c->wsum); cudaFree(c->Bt32);
    c->eta = c->alpha = c->Elt = c->Bt = c->mw = c->phi = c->phi_KV = c->kbuf = c->alpha_ss = c->e_dead = c->wsum = nullptr;
    c->Bt32 = nullptr;
    c->model_set = false;
}

// (LK, J) lane shape for K topics: smallest padded width among the compiled shapes.
// v2 adds (4, 13) (K = 100 -> 104 padded columns instead of 112); PYLDA_SHAPE="LK,J" forces a shape.
bool pick_shape(int K, int* LK, int* J, bool v2 = false, int W = 8) {
    const int pairs = (K + 1) / 2;
    if (v2) {
        const char* env = getenv("PYLDA_SHAPE");
        int lk = 0, j = 0;
        if (env && sscanf(env, "%d,%d", &lk, &j) == 2 && lk * j >= pairs &&
            (((lk == 1 || lk == 2 || lk == 4 || lk == 8 || lk == 16 || lk == 32) && (j == 5 || j == 7 || j == 8)) ||
             (lk == 4 && j == 13) || (lk == 32 && j == 16))) {
            *LK = lk; *J = j;
            return true;
        }
    }
    const int Js[] = {5, 7, 8};
    int best = 1 << 30;
    bool found = false;
    for (int lk = 1; lk <= 32; lk <<= 1) {
        for (int j : Js) {
            if (lk * j >= pairs && lk * j < best) { best = lk * j; *LK = lk; *J = j; found = true; }
        }
    }
    // (4, 13): fewer shuffle levels and less padding, measured faster for groups of one or two warps
    // (short documents); the wider (8, 7) keeps the per-lane register footprint of long rows smaller
    if (v2 && W <= 2 && 4 * 13 >= pairs && 4 * 13 < best) { best = 52; *LK = 4; *J = 13; found = true; }
    if (!found && 32 * 16 >= pairs) { *LK = 32; *J = 16; found = true; }
    return found;
}

const void* lookup_v2(int LK, int J, int W, int V) {
    switch (LK) {
        case 1: return estep_v2_lk1(J, W, V);
        case 2: return estep_v2_lk2(J, W, V);
        case 4: return estep_v2_lk4(J, W, V);
        case 8: return estep_v2_lk8(J, W, V);
        case 16: return estep_v2_lk16(J, W, V);
        case 32: return estep_v2_lk32(J, W, V);
    }
    return nullptr;
}

const void* lookup_rt(int LK, int J, int W, int Rsel, int NW, int* R) {
    switch (LK) {
        case 1: return estep_rt_lk1(J, W, Rsel, NW, R);
        case 2: return estep_rt_lk2(J, W, Rsel, NW, R);
        case 4: return estep_rt_lk4(J, W, Rsel, NW, R);
        case 8: return estep_rt_lk8(J, W, Rsel, NW, R);
        case 16: return estep_rt_lk16(J, W, Rsel, NW, R);
        case 32: return estep_rt_lk32(J, W, Rsel, NW, R);
    }
    return nullptr;
}

// bank-conflict-free shared-memory row stride (doubles) for LDS.128 with LK lanes per row
int tile_stride(int KP, int LK) {
    int st = KP;
    if (LK == 1) { while (st % 4 != 2) st += 2; }
    else if (LK == 2) { while (st % 8 != 4) st += 2; }
    else if (LK == 4) { while (st % 16 != 8) st += 2; }
    return st;
}

struct GroupLayout {
    int off_gam, off_spart, off_red, off_cnt, off_mwr, off_rid, off_tile, bytes;
};
int align_up(int x, int a) { return (x + a - 1) / a * a; }
// shared-memory layout of one document group of estep_v2 (W warps)
GroupLayout group_layout_v2(int W, int LK, int KPAD, int nmax, int ST) {
    GroupLayout g;
    const int LN = 32 / LK;
    const int NP = (W >= 4) ? W : W * LN;
    int o = 16;                       // mbarrier + queue slot
    o += KPAD * 8;                    // es
    g.off_gam = 0;
    g.off_spart = o; o += NP * KPAD * 8;
    g.off_red = o;   o += align_up(3 * W + 2, 2) * 8;
    g.off_cnt = o;   o += nmax * 8;
    g.off_mwr = o;   o += nmax * 8;
    g.off_rid = o;   o += nmax * 4;
    o = align_up(o, 16);
    g.off_tile = o;  o += nmax * ST * 8 + KPAD * 8;   // + slack for the unpredicated over-read of the last row
    g.bytes = align_up(o, 128);
    return g;
}

// shared-memory layout of one document group of estep_rt (W warps, cap = W * LN * R rows)
GroupLayout group_layout_rt(int W, int LK, int KPAD, int cap, int ST) {
    GroupLayout g;
    const int LN = 32 / LK;
    const int WL = W * LN;
    const int NB = WL > 128 ? 3 : WL > 64 ? 2 : WL > 32 ? 1 : 0;
    const int NP = WL >> NB;
    int o = 16;                       // mbarrier + queue slot
    o += 2 * KPAD * 8;                // e, double buffered
    g.off_gam = 0;
    g.off_spart = o; o += NP * KPAD * 8;
    g.off_red = o;   o += align_up(4 * W + 2, 2) * 8;   // dsum [W], ELBO pairs [2W], live counts [W]
    g.off_cnt = o;   o += cap * 8;
    g.off_mwr = o;   o += cap * 8;
    g.off_rid = o;   o += cap * 4;
    o = align_up(o, 16);
    g.off_tile = o;  o += cap * ST * 8 + KPAD * 8;
    g.bytes = align_up(o, 128);
    return g;
}

// shared-memory layout of one CTA of estep_hy (8 warps; `cap` rows in shared memory, `capr` in registers,
// C CTAs per cluster)
GroupLayout group_layout_hy(int LK, int KPAD, int C, int capr, int cap, int ST) {
    GroupLayout g;
    const int W = 8, LN = 32 / LK;
    int o = 16;
    o += 2 * KPAD * 8;                // e, double buffered
    g.off_spart = o; o += W * LN * KPAD * 8;
    g.off_red = o;   o += align_up(4 * W + 2, 2) * 8;
    g.off_gam = o;   o += (C > 1 ? 2 * C * KPAD * 8 : 0);   // exchange slots [2][C][KPAD]
    const int rows = capr + cap + LN;
    g.off_cnt = o;   o += rows * 8;
    g.off_mwr = o;   o += rows * 8;
    g.off_rid = o;   o += rows * 4;
    o = align_up(o, 16);
    g.off_tile = o;  o += std::max(cap, capr) * ST * 8 + KPAD * 8;
    g.bytes = align_up(o, 128);
    return g;
}

int ensure_partial(pylda_ctx* ctx, size_t n) {
    if (ctx->partial_cap >= n) return 0;
    CK(dalloc(&ctx->partial, n));
    ctx->partial_cap = n;
    return 0;
}

// psisum[k] = psi(sum_v eta_kv), rowsum[k] = sum_v eta_kv (scratch when the caller does not need it)
int rowsum_psi(pylda_ctx* ctx, const double* eta, int K, int V, double* psisum, double* rowsum) {
    const int chunks = std::max(1, std::min(64, (4 * ctx->prop.multiProcessorCount + K - 1) / K));
    if (ensure_partial(ctx, (size_t)K * chunks)) return 1;
    k_rowsum<<<dim3(K, chunks), 256, 0, ctx->stream>>>(eta, K, V, ctx->partial);
    k_psi_of_rowsum<<<(K + 127) / 128, 128, 0, ctx->stream>>>(ctx->partial, K, chunks, psisum, rowsum);
    return 0;
}

int prepare_tables(pylda_ctx* ctx, bool heldout, int* launches) {
    const int K = ctx->K, V = ctx->V, KP = ctx->KP;
    double* psisum = ctx->kbuf;
    if (rowsum_psi(ctx, ctx->eta, K, V, psisum, ctx->kbuf + K)) return 1;
    dim3 tb(32, 8), tg((V + 31) / 32, (K + 31) / 32);
    k_elog_transpose<<<tg, tb, 0, ctx->stream>>>(ctx->eta, psisum, K, V, KP, ctx->Elt);
    const int nb = std::min((V + 7) / 8, ctx->prop.multiProcessorCount * 8);
    k_build_B<<<nb, 256, 0, ctx->stream>>>(ctx->Elt, K, V, KP, ctx->Bt, ctx->mw, ctx->phi, ctx->flat_part);
    k_reduce_final<<<1, 256, 0, ctx->stream>>>(ctx->flat_part, nb, 1, ctx->flat_dev);
    CK(cudaMemcpyAsync(ctx->flat_host, ctx->flat_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->ev_flat, ctx->stream));
    *launches += 5;
    if (heldout) {
        const int nchunk = 64;
        if (ensure_partial(ctx, (size_t)2 * nchunk * K)) return 1;
        dim3 lg((K + 31) / 32, nchunk);
        k_lse_partial<<<lg, 256, 0, ctx->stream>>>(ctx->Elt, K, V, KP, ctx->partial, ctx->partial + (size_t)nchunk * K);
        k_lse_final<<<(K + 127) / 128, 128, 0, ctx->stream>>>(ctx->partial, ctx->partial + (size_t)nchunk * K, K, nchunk,
                                                             ctx->kbuf + 2 * K);
        *launches += 2;
    }
    CK(cudaGetLastError());
    return 0;
}

// Length classes of the single-CTA resident paths.  Documents are sorted by n_d (descending); every
// class has a row capacity and a document goes to the class with the smallest capacity that holds it.
//   "8x1"  estep_v2: tile in shared memory, 8 warps per document, one document per CTA;
//   "rW"   estep_rt: tile in registers, W warps per document, 8/W documents in flight per CTA.
// PYLDA_CLASSES overrides the default list (tuning aid).
struct ClassCfg { int kind, W, G, R, LK, J; };

std::vector<ClassCfg> class_config(bool use_hy) {
    std::vector<ClassCfg> out;
    const char* env = getenv("PYLDA_CLASSES");
    // with the hybrid kernel every document above the largest register-tile class is its business
    std::string spec = env && *env ? env : use_hy ? "r8,r4,r2,r1" : "8x1,r8,r4,r2,r1";
    size_t pos = 0;
    while (pos < spec.size()) {
        size_t end = spec.find(',', pos);
        if (end == std::string::npos) end = spec.size();
        const std::string item = spec.substr(pos, end - pos);
        int W = 0;
        if (sscanf(item.c_str(), "r%d", &W) == 1) {
            if (W == 1 || W == 2 || W == 4 || W == 8) out.push_back({1, W, 8 / W, 0, 0, 0});
        } else if (item == "8x1") {
            out.push_back({0, 8, 1, 0, 0, 0});
        }
        pos = end + 1;
    }
    if (out.empty()) out = {{0, 8, 1, 0, 0, 0}};
    return out;
}

// PYLDA_PROFILE_CLASSES=1: CUDA-event time of every class launch, printed to stderr (tuning aid)
struct ClassTimer {
    struct Rec { cudaEvent_t a, b; std::string tag; };
    std::vector<Rec> recs;
    bool on = getenv("PYLDA_PROFILE_CLASSES") != nullptr;
    void begin(cudaStream_t s, const char* fmt, ...) {
        if (!on) return;
        char buf[256];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        Rec r; cudaEventCreate(&r.a); cudaEventCreate(&r.b); r.tag = buf;
        cudaEventRecord(r.a, s);
        recs.push_back(r);
    }
    void end(cudaStream_t s) { if (on) cudaEventRecord(recs.back().b, s); }
    void report(cudaStream_t s) {
        if (!on) return;
        cudaStreamSynchronize(s);
        for (auto& r : recs) {
            float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
            fprintf(stderr, "[pylda class] %-60s %9.3f ms\n", r.tag.c_str(), ms);
            cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        }
        recs.clear();
    }
};

// second-generation streaming kernel for documents [lo, hi) of the sorted order (all with n <= nmax)
int launch_stream2(pylda_ctx* ctx, Corpus& cp, long long lo, long long hi, int nmax, int LK, int J, int max_iter,
                   double tol, pylda_stats* st, int counter_slot, int* launched, int park_nc, int park_long) {
    *launched = 0;
    const int K = ctx->K, KP = ctx->KP;
    const int KPAD = 2 * LK * J, LN = 32 / LK, W = 8;
    // rows of ids / counts held in shared memory at a time: the longest document of the class, but never more than
    // keeps two CTAs per SM (longer documents are walked in chunks)
    const int fixed = 16 + KPAD * 8 + W * KPAD * 8 + align_up(3 * W + 2, 2) * 8 + 256;
    const int limit = (((int)ctx->prop.sharedMemPerBlockOptin / 2 - fixed) / 12) / (W * LN) * (W * LN);
    if (limit < W * LN) return fail(ctx, "internal: no shared memory left for the streaming kernel (K=%d)", K);
    const int cap = std::min((nmax + W * LN - 1) / (W * LN) * (W * LN), limit);
    // lean instantiation when no document of the class can be handed over (all longer than 192 terms) or chunked
    const int nmin = cp.n_sorted[(size_t)hi - 1];
    const bool parks = (park_nc > 0 && nmin <= 192) || (park_long > 192 && nmin <= park_long);
    const void* fn = estep_stream_lookup(LK, J, nmax > cap ? 1 : parks ? 2 : 0);
    if (!fn) return 0;
    GroupLayout gl;
    int o = 16 + KPAD * 8;
    gl.off_spart = o; o += W * KPAD * 8;
    gl.off_red = o;   o += align_up(3 * W + 2, 2) * 8;
    gl.off_cnt = o;   o += cap * 8;
    gl.off_rid = o;   o += cap * 4;
    gl.bytes = align_up(o, 128);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, gl.bytes));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&occ, fn, 256, gl.bytes, 0));
    if (occ < 1) return 0;
    const long long nd = hi - lo;
    const long long grid = std::min<long long>((long long)ctx->prop.multiProcessorCount * occ, nd);
    EParams p;
    memset(&p, 0, sizeof p);
    p.row_ptr = cp.row_ptr; p.ids = cp.ids; p.cts = cp.cts;
    p.order = cp.order + lo; p.ndocs = (int)nd; p.counter = ctx->counters + counter_slot;
    p.Bt = ctx->Bt; p.mw = ctx->mw; p.alpha = ctx->alpha; p.alpha_max = ctx->alpha_max;
    p.gamma = cp.gamma_dst; p.phi_ss = ctx->phi; p.docterm = cp.docterm; p.iters = cp.iters;
    p.K = K; p.KP = KP; p.ST = KP; p.max_iter = max_iter; p.tol = tol;
    p.W = W; p.nmax = cap; p.group_bytes = gl.bytes;
    {
        const char* z = getenv("PYLDA_ZIGZAG");                 // alternate the row order trip by trip (L1 reuse)
        p.compact = !(z && !strcmp(z, "0"));
    }
    p.park_nc = park_nc; p.park_long = park_long;
    p.park_rec = cp.park_rec; p.park_gam = cp.park_gam; p.park_lists = cp.park_lists;
    p.park_counts = ctx->park_ctr; p.park_cap = (int)cp.D;
    p.off_spart = gl.off_spart; p.off_red = gl.off_red; p.off_cnt = gl.off_cnt; p.off_rid = gl.off_rid;
    void* args[] = {&p};
    CK(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(256), args, (size_t)gl.bytes, ctx->stream));
    st->n_launches++;
    st->n_estep_launches++;
    *launched = 1;
    return 0;
}

// psi(x), x > 0, on the host (recurrence + asymptotic series; ~1e-13): only feeds the safety bound below
double host_digamma(double x) {
    double r = 0.0;
    while (x < 8.0) { r -= 1.0 / x; x += 1.0; }
    const double u = 1.0 / (x * x);
    return r + log(x) - 0.5 / x - u * (1.0 / 12 - u * (1.0 / 120 - u * (1.0 / 252 - u * (1.0 / 240 - u * (1.0 / 132)))));
}

// Narrow stages (estep_narrow.cuh): configuration of one E-step.  A topic eliminated as dead stays dead while
// e_k s_k < ulp(alpha_k)/2, s_k <= sum_n w_n; chk_bound = alpha_min 2^-54 / exp(psi(alpha_max)) is the value
// of sum_n w_n below which that is certain.  The hand-over is only used when the bound leaves three orders
// of magnitude of room (alpha <~ 0.02); the kernels check every document against it.
struct ParkCfg { int nc; double chk_bound; int long_rows; };   // long_rows: longest document estep_longc takes (0: off)
ParkCfg park_config(const pylda_ctx* ctx) {
    ParkCfg c;
    const char* e = getenv("PYLDA_PARK");
    c.nc = 16;
    if (e && *e) c.nc = atoi(e);
    if (c.nc != 8 && c.nc != 16) c.nc = 0;
    const double ed = exp(host_digamma(ctx->alpha_max));
    c.chk_bound = ed > 0.0 ? ctx->alpha_min * ldexp(1.0, -54) / ed : HUGE_VAL;
    if (!(c.chk_bound >= 1e3)) c.nc = 0;
    const char* ce = getenv("PYLDA_COMPACT");
    if ((ce && !strcmp(ce, "0")) || ctx->force_full) c.nc = 0;
    // Long documents (> 192 terms) are handed to estep_longc at <= 32 live topics; its per-CTA scratch is sized for
    // the longest document it may get, hence the cap.
    c.long_rows = (c.nc >= 16) ? 16384 : 0;
    if (const char* pl = getenv("PYLDA_PARK_LONG")) c.long_rows = (c.nc >= 16) ? std::max(0, atoi(pl)) : 0;
    return c;
}

// parameters common to the narrow stages and estep_longc, for park list li
NParams narrow_params(pylda_ctx* ctx, Corpus& cp, const ParkCfg& pc, int li, int max_iter, double tol) {
    NParams p;
    memset(&p, 0, sizeof p);
    p.row_ptr = cp.row_ptr; p.ids = cp.ids; p.cts = cp.cts;
    p.Bt = ctx->Bt; p.mw = ctx->mw; p.alpha = ctx->alpha; p.e_dead = ctx->e_dead;
    p.gamma = cp.gamma_dst; p.phi_ss = ctx->phi; p.wsum = ctx->wsum; p.docterm = cp.docterm; p.iters = cp.iters;
    p.K = ctx->K; p.KP = ctx->KP; p.max_iter = max_iter; p.tol = tol;
    p.lg_alpha = ctx->lg_alpha; p.alpha_sum = ctx->alpha_sum;
    p.list = cp.park_lists + (size_t)li * cp.D; p.count = ctx->park_ctr + li; p.head = ctx->park_ctr + 16 + li;
    p.rec = cp.park_rec; p.gam = cp.park_gam; p.lists = cp.park_lists; p.counts = ctx->park_ctr; p.cap = (int)cp.D;
    p.chk_bound = pc.chk_bound; p.revived = ctx->counters + 14;
    return p;
}

// before the first kernel that adds to wsum
int narrow_begin(pylda_ctx* ctx, pylda_stats* st) {
    k_e_dead<<<(ctx->K + 127) / 128, 128, 0, ctx->stream>>>(ctx->alpha, ctx->K, ctx->e_dead);
    CK(cudaMemsetAsync(ctx->wsum, 0, (size_t)ctx->V * sizeof(double), ctx->stream));
    st->n_launches++;
    return 0;
}

// after the last one: the statistics of the eliminated topics, for every document that finished in a compact stage
int narrow_end(pylda_ctx* ctx, pylda_stats* st, ClassTimer& timer) {
    timer.begin(ctx->stream, "dead-topic statistics (k_dead_phi)");
    k_dead_phi<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(ctx->Bt, ctx->wsum, ctx->e_dead, ctx->K, ctx->V,
                                                                         ctx->KP, ctx->phi);
    timer.end(ctx->stream);
    CK(cudaGetLastError());
    st->n_launches++;
    return 0;
}

// Compact stage for long documents (estep_longc.cuh): park list 9, fed by estep_stream and estep_v2.
int launch_longc(pylda_ctx* ctx, Corpus& cp, const ParkCfg& pc, int nmax, long long ndocs_max, int max_iter, double tol,
                 pylda_stats* st, ClassTimer& timer) {
    int want_occ = 2;
    if (const char* e = getenv("PYLDA_LONGC_CTAS")) want_occ = (atoi(e) == 3) ? 3 : 2;
    const void* fn = estep_longc_lookup(32, want_occ);
    if (!fn) return fail(ctx, "no compact-stage kernel for long documents");
    const int NC = 32, K = ctx->K;
    const int fixed = (NC + 8 * NC + 4 * 8 + 2 + ((K + 1) & ~1)) * (int)sizeof(double);
    // shared-memory tile: documents of up to smem_rows rows never leave the SM; two (three) CTAs per SM
    int budget = (want_occ == 3 ? 72 : 96) * 1024;
    if (const char* e = getenv("PYLDA_LONGC_SMEM")) budget = std::max(0, atoi(e)) * 1024;
    int smem_rows = std::max(0, (budget - fixed) / (NC * 8)) / 64 * 64;
    const int scratch_rows = std::min((nmax + 63) / 64 * 64, (pc.long_rows + 63) / 64 * 64);
    smem_rows = std::min(smem_rows, scratch_rows);
    const int smem = fixed + smem_rows * NC * 8;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&occ, fn, 256, smem, 0));
    if (occ < 1) return fail(ctx, "internal: zero occupancy for the long-document compact stage (smem %d)", smem);
    if (const char* e = getenv("PYLDA_LONGC_OCC")) occ = std::max(1, std::min(occ, atoi(e)));
    const long long grid = std::max<long long>(1, std::min<long long>((long long)ctx->prop.multiProcessorCount * occ, ndocs_max));
    const size_t need = (size_t)grid * scratch_rows;
    if (ctx->longc_rows < need) {
        if (ctx->longc_tile) { cudaFree(ctx->longc_tile); cudaFree(ctx->longc_cnt); cudaFree(ctx->longc_ids); }
        ctx->longc_tile = nullptr; ctx->longc_cnt = nullptr; ctx->longc_ids = nullptr; ctx->longc_rows = 0;
        CK(dalloc(&ctx->longc_tile, need * NC));
        CK(dalloc(&ctx->longc_cnt, need));
        CK(dalloc(&ctx->longc_ids, need));
        ctx->longc_rows = need;
    }
    LParams lp;
    lp.n = narrow_params(ctx, cp, pc, 9, max_iter, tol);
    lp.scratch_tile = ctx->longc_tile; lp.scratch_cnt = ctx->longc_cnt; lp.scratch_ids = ctx->longc_ids;
    lp.scratch_rows = scratch_rows; lp.smem_rows = smem_rows;
    lp.mix = 4;
    if (const char* e = getenv("PYLDA_LONGC_MIX")) lp.mix = std::max(1, atoi(e));
    void* args[] = {&lp};
    timer.begin(ctx->stream, "longc<%d> smem=%d (tile rows %d) scratch rows %d grid=%lld", NC, smem, smem_rows, scratch_rows, grid);
    CK(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(256), args, (size_t)smem, ctx->stream));
    timer.end(ctx->stream);
    st->n_launches++;
    st->n_estep_launches++;
    return 0;
}

int launch_narrow(pylda_ctx* ctx, Corpus& cp, const ParkCfg& pc, int max_iter, double tol, pylda_stats* st, ClassTimer& timer) {
    // launch order: the 32-column list (8) feeds the 16-column ones (0, 1, 2, 7), which feed the 8-column ones (3..6)
    static const int order[PARK_NARROW_LISTS] = {8, 0, 1, 2, 7, 3, 4, 5, 6};
    static const int NCs[PARK_NARROW_LISTS] = {16, 16, 16, 8, 8, 8, 8, 16, 32};
    static const int Gs[PARK_NARROW_LISTS] = {8, 16, 32, 4, 8, 16, 32, 32, 32};
    static const int RPLs[PARK_NARROW_LISTS] = {3, 3, 3, 6, 6, 6, 6, 6, 3};
    for (int oi = 0; oi < PARK_NARROW_LISTS; ++oi) {
        const int li = order[oi];
        if (NCs[li] > 8 && pc.nc < 16) continue;        // PYLDA_PARK=8: the 8-column stage only
        const int NC = NCs[li], G = Gs[li];
        // CTAs per SM the kernel is compiled for: 2 (255 registers, no spills) or 3 (168 registers, tuning aid)
        int minb = 2;
        if (const char* e = getenv("PYLDA_NARROW_OCC")) minb = (atoi(e) == 3 && RPLs[li] * NC <= 48) ? 3 : 2;
        const void* fn = estep_narrow_lookup(NC, G, RPLs[li], minb);
        if (!fn) return fail(ctx, "no narrow-stage kernel for NC=%d G=%d", NC, G);
        const int smem = 4 * (32 * (NC + 2) + (32 / G) * NC + (32 / G) * 16) * (int)sizeof(double);
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&occ, fn, 128, smem, 0));
        if (occ < 1) return fail(ctx, "internal: zero occupancy for the narrow stage NC=%d G=%d", NC, G);
        // the list length is only known on the device: size the grid for the documents that could be there
        long long grid = (long long)ctx->prop.multiProcessorCount * occ;
        grid = std::max<long long>(1, std::min<long long>(grid, (cp.D * G + 127) / 128));
        NParams p = narrow_params(ctx, cp, pc, li, max_iter, tol);
        {
            // which stages hand over (bit 0: 32 -> 16 columns, bit 1: 16 -> 8)
            const char* h = getenv("PYLDA_NARROW_HANDOVER");
            const int mask = h ? atoi(h) : 3;
            p.handover = (NC == 32) ? (mask & 1) : (NC == 16) ? ((mask >> 1) & 1) : 0;
        }
        void* args[] = {&p};
        timer.begin(ctx->stream, "narrow<%d,%d,%d> smem=%d grid=%lld", NC, G, RPLs[li], smem, grid);
        CK(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(128), args, (size_t)smem, ctx->stream));
        timer.end(ctx->stream);
        st->n_launches++;
        st->n_estep_launches++;
    }
    return 0;
}

// Precision sweep (BASELINE.json configs[3]): PYLDA_PRECISION = f64s | tile32 | f32 runs estep_sweep.cuh instead of
// the product kernels (measurement tool; K <= 256).
int launch_estep_sweep(pylda_ctx* ctx, Corpus& cp, const char* mode, int max_iter, double tol, pylda_stats* st) {
    const int K = ctx->K, KP = ctx->KP;
    if (K > 256) return fail(ctx, "PYLDA_PRECISION=%s: the sweep kernel handles K <= 256 (got %d)", mode, K);
    EParams p;
    memset(&p, 0, sizeof p);
    p.row_ptr = cp.row_ptr; p.ids = cp.ids; p.cts = cp.cts;
    p.order = cp.order; p.ndocs = (int)cp.D;
    p.Bt = ctx->Bt; p.mw = ctx->mw; p.alpha = ctx->alpha;
    p.gamma = cp.gamma_dst; p.phi_ss = ctx->phi; p.docterm = cp.docterm; p.iters = cp.iters;
    p.K = K; p.KP = KP; p.max_iter = max_iter; p.tol = tol;
    const int grid = ctx->prop.multiProcessorCount * 8;
    if (!strcmp(mode, "f64s")) {
        estep_sweep<double, double><<<grid, 256, 0, ctx->stream>>>(p, ctx->Bt);
    } else {
        const size_t n = (size_t)ctx->V * KP;
        if (!ctx->Bt32) CK(cudaMalloc((void**)&ctx->Bt32, n * sizeof(float)));
        k_convert_table<float><<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(ctx->Bt, n, ctx->Bt32);
        if (!strcmp(mode, "tile32")) estep_sweep<float, double><<<grid, 256, 0, ctx->stream>>>(p, ctx->Bt32);
        else if (!strcmp(mode, "f32")) estep_sweep<float, float><<<grid, 256, 0, ctx->stream>>>(p, ctx->Bt32);
        else return fail(ctx, "PYLDA_PRECISION must be f64s, tile32 or f32 (got %s)", mode);
        st->n_launches++;
    }
    CK(cudaGetLastError());
    st->n_launches++;
    st->n_estep_launches++;
    st->docs_streamed = cp.D;
    return 0;
}

int launch_estep(pylda_ctx* ctx, Corpus& cp, int max_iter, double tol, pylda_stats* st) {
    const int K = ctx->K, KP = ctx->KP;
    int LK = 0, J = 0;
    if (!pick_shape(K, &LK, &J, true, 8))
        return fail(ctx, "unsupported number of topics K=%d (max 1024)", K);
    const int KPAD = 2 * LK * J;                           // shape of the long-document paths (cluster)
    const int LN = 32 / LK;
    const int ST = KP;                                     // rows packed (LDS.128 over LK lanes is conflict-free per quarter warp)
    const int smem_budget = (int)ctx->prop.sharedMemPerBlockOptin;
    const char* kv = getenv("PYLDA_KERNEL");
    const bool use_rt = !(kv && !strcmp(kv, "v2"));
    int R_hy = 0;
    // The hybrid register / shared-memory cluster kernel keeps whole long documents on chip (PYLDA_KERNEL=hybrid).
    // Measured (profiles/r2b_*): its per-trip serial phase (owner sums, exp(psi), cluster exchange) costs what the
    // residency saves -- 73 ms vs 64 ms for the long documents of the headline config -- so streaming stays the default.
    const bool want_hy = kv && (!strcmp(kv, "hybrid") || !strcmp(kv, "cluster"));
    const void* fn_hy = want_hy ? estep_hy_lookup(LK, J, &R_hy) : nullptr;

    // Candidate classes, each with a row capacity; a document goes to the class with the smallest
    // capacity that holds it.  kind 0 = estep_v2 (tile in shared memory), kind 1 = estep_rt (tile
    // in registers); documents above every capacity use the cluster / streaming kernels.  The lane
    // shape (LK, J) is chosen per class.
    struct Cls { int kind, W, G, cap, LK, J; const void* fn; long long lo, hi; };
    std::vector<Cls> cls;
    for (const ClassCfg& c : class_config(fn_hy != nullptr)) {
        int lk = c.LK, j = c.J;
        if (lk > 0) {
            if (lk * j < (K + 1) / 2) continue;            // forced shape too narrow for K
        } else if (!pick_shape(K, &lk, &j, true, c.W)) {
            continue;
        }
        const int kpad = 2 * lk * j, ln = 32 / lk;
        if (kpad > 128 * c.W) continue;                    // at most 4 topics per owner thread
        if (c.kind == 0) {
            const int avail = (smem_budget / c.G) & ~127;
            const GroupLayout g0 = group_layout_v2(c.W, lk, kpad, 0, ST);
            int n = (avail - g0.bytes - 128) / (ST * 8 + 20);
            n = n / ln * ln;
            const void* fn = lookup_v2(lk, j, c.W, 0);
            if (n > 0 && fn) cls.push_back({0, c.W, c.G, n, lk, j, fn, 0, 0});
        } else if (use_rt) {
            int R = 0;
            const void* fn = lookup_rt(lk, j, c.W, c.R, c.R ? c.W * c.G : 0, &R);
            if (!fn) continue;
            const int cap = c.W * ln * R;
            const GroupLayout gl = group_layout_rt(c.W, lk, kpad, cap, ST);
            if (c.G * gl.bytes + align_up(kpad * 8, 128) > smem_budget) continue;
            cls.push_back({1, c.W, c.G, cap, lk, j, fn, 0, 0});
        }
    }
    std::sort(cls.begin(), cls.end(), [](const Cls& a, const Cls& b) { return a.cap > b.cap; });
    // equal capacities: keep the first
    cls.erase(std::unique(cls.begin(), cls.end(), [](const Cls& a, const Cls& b) { return a.cap == b.cap; }), cls.end());
    const int NC = (int)cls.size();
    if (NC > 12) return fail(ctx, "too many length classes");

    const std::vector<int>& ns = cp.n_sorted;
    const long long D = cp.D;
    auto first_leq = [&](int limit) -> long long {   // first index (descending order) whose n <= limit
        return std::partition_point(ns.begin(), ns.end(), [&](int n) { return n > limit; }) - ns.begin();
    };
    CK(cudaMemsetAsync(ctx->counters, 0, 32 * sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(ctx->park_ctr, 0, PARK_CTRS * sizeof(int), ctx->stream));
    const ParkCfg pc = park_config(ctx);
    for (int i = 0; i < NC; ++i) cls[i].lo = first_leq(cls[i].cap);
    for (int i = 0; i < NC; ++i) cls[i].hi = (i + 1 < NC) ? cls[i + 1].lo : D;
    long long nlong = NC ? cls[0].lo : D;          // documents [0, nlong) fit no single-CTA class
    ClassTimer timer;
    long long nstream = nlong;
    if (nlong > 0 && fn_hy) {
        // hybrid register / shared-memory tile kernel: clusters of C = 1, 2, 4, 8 CTAs, carved from the short end
        const int capr = 8 * LN * R_hy;
        bool zeroed = false;
        long long hi = nlong;
        for (int C = 1; C <= 8 && hi > 0; C <<= 1) {
            const GroupLayout g0 = group_layout_hy(LK, KPAD, C, 0, 0, ST);
            int cap_s = (smem_budget - g0.bytes - 256 - (capr + LN) * 20) / (ST * 8 + 20);
            cap_s = cap_s / LN * LN;
            if (cap_s < capr) break;                    // the register rows leave through the shared-memory region
            const int per_cta = capr + cap_s - LN;      // (the per-CTA slice is rounded up to LN rows)
            const long long lo = first_leq(C * per_cta);
            if (lo >= hi) continue;
            const long long nd = hi - lo;
            const GroupLayout gl = group_layout_hy(LK, KPAD, C, capr, cap_s, ST);
            CK(cudaFuncSetAttribute(fn_hy, cudaFuncAttributeMaxDynamicSharedMemorySize, gl.bytes));
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = (size_t)gl.bytes;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = C; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            cfg.gridDim = dim3((unsigned)(C * ctx->prop.multiProcessorCount));
            int ncl = 0;
            if (C > 1) {
                CK(cudaOccupancyMaxActiveClusters(&ncl, fn_hy, &cfg));
                if (ncl < 1) break;                     // this cluster size cannot be scheduled: the rest streams
            } else {
                ncl = ctx->prop.multiProcessorCount;
            }
            ncl = (int)std::min<long long>(ncl, nd);
            cfg.gridDim = dim3((unsigned)(ncl * C));
            if (!zeroed) {
                CK(cudaMemsetAsync(cp.docterm, 0, (size_t)std::max<long long>(D, 1) * sizeof(double), ctx->stream));
                zeroed = true;
            }
            EParams p;
            memset(&p, 0, sizeof p);
            p.row_ptr = cp.row_ptr; p.ids = cp.ids; p.cts = cp.cts;
            p.order = cp.order + lo; p.ndocs = (int)nd; p.counter = nullptr;
            p.Bt = ctx->Bt; p.mw = ctx->mw; p.alpha = ctx->alpha; p.alpha_max = ctx->alpha_max;
            p.gamma = cp.gamma_dst; p.phi_ss = ctx->phi; p.docterm = cp.docterm; p.iters = cp.iters;
            p.K = K; p.KP = KP; p.ST = ST; p.max_iter = max_iter; p.tol = tol;
            p.W = 8; p.nmax = cap_s; p.group_bytes = gl.bytes; p.off_groups = 0;
            p.off_gam = gl.off_gam; p.off_spart = gl.off_spart; p.off_red = gl.off_red; p.off_cnt = gl.off_cnt;
            p.off_mwr = gl.off_mwr; p.off_rid = gl.off_rid; p.off_tile = gl.off_tile;
            void* args[] = {&p};
            timer.begin(ctx->stream, "hybrid<%d,%d,R%d> C=%d docs=%lld nmax=%d nmin=%d rows/CTA=%d+%d smem=%d clusters=%d", LK, J,
                        R_hy, C, nd, ns[lo], ns[hi - 1], capr, cap_s, gl.bytes, ncl);
            CK(cudaLaunchKernelExC(&cfg, fn_hy, args));
            timer.end(ctx->stream);
            st->n_launches++;
            st->n_estep_launches++;
            hi = lo;
        }
        nstream = hi;
    }
    st->docs_streamed = nstream;
    st->docs_resident = D - nstream;
    // Hand-over of LONG documents from the streaming kernel: its parking instantiation costs ~6 % in the trip loop
    // (register spills), which pays when long documents get down to 32 live topics with many trips left -- any model
    // after the first M-step (trip ~20 of 50) -- and does not at EM iteration 1, where the random initial topics are
    // nearly flat and a long document keeps more than 32 topics alive until trip ~45.  The two are told apart by the
    // model itself: mean_w sum_k B[w,k] / K (B = exp(E[log beta]) scaled to max 1 per word; k_build_B) is ~0.7 for
    // eta0 ~ Gamma(100, 1/100) and < 0.1 after one M-step.  A deterministic function of eta: the same model always
    // takes the same path.  Only speed depends on it.
    int stream_long = pc.long_rows;
    if (stream_long > 0 && nstream > 0) {
        CK(cudaEventSynchronize(ctx->ev_flat));
        double thr = 0.3;
        if (const char* e = getenv("PYLDA_FLAT_THRESHOLD")) thr = atof(e);
        if (*ctx->flat_host / ((double)ctx->V * K) > thr) stream_long = 0;
    }
    auto launch_long = [&]() -> int {
        // Long documents: the streaming kernel.  (Tried in round 2 and dropped, both measured slower at the headline
        // config -- DESIGN.md section 6: a cluster kernel with the whole tile on chip [estep_hy, kept opt-in], and a
        // 16-warp streaming kernel with a resident shared-memory prefix: streaming is bound by the bytes it keeps in
        // flight towards L2, and one document per SM keeps fewer in flight than two.)
        if (nstream <= 0) return 0;
        int launched = 0;
        timer.begin(ctx->stream, "stream2<%d,%d> docs=%lld nmax=%d", LK, J, nstream, ns[0]);
        if (launch_stream2(ctx, cp, 0, nstream, ns[0], LK, J, max_iter, tol, st, 17, &launched, pc.nc, stream_long)) return 1;
        timer.end(ctx->stream);
        if (!launched) return fail(ctx, "no streaming kernel instantiation for LK=%d J=%d", LK, J);
        return 0;
    };
    auto launch_class = [&](int ci) -> int {
        const Cls& c = cls[ci];
        const long long nd = c.hi - c.lo;
        if (nd <= 0) return 0;
        const int W = c.W;
        const int kpad = 2 * c.LK * c.J, ln = 32 / c.LK;
        int nmax, G;
        GroupLayout gl;
        if (c.kind == 1) {
            nmax = c.cap;
            gl = group_layout_rt(W, c.LK, kpad, c.cap, ST);
            G = c.G;
            if (const char* e = getenv("PYLDA_RT_G")) G = std::max(1, std::min(G, atoi(e)));   // tuning aid
        } else {
            nmax = std::max(ln, (ns[c.lo] + ln - 1) / ln * ln);
            gl = group_layout_v2(W, c.LK, kpad, nmax, ST);
            // shorter documents than the class limit: pack more groups per CTA, up to the thread bound
            G = std::min(c.G, smem_budget / gl.bytes);
        }
        if (G < 1) return fail(ctx, "internal: class %d needs %d B of shared memory per group", ci, gl.bytes);
        G = (int)std::min<long long>(G, nd);
        const int smem = G * gl.bytes + (c.kind == 1 ? align_up(kpad * 8, 128) : 0);   // rt: + the CTA's alpha copy
        const int threads = G * W * 32;
        CK(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&occ, c.fn, threads, smem, 0));
        if (occ < 1) return fail(ctx, "internal: zero occupancy for class %d (smem %d, threads %d)", ci, smem, threads);
        long long grid = (long long)ctx->prop.multiProcessorCount * occ;
        grid = std::min(grid, (nd + G - 1) / G);
        EParams p;
        memset(&p, 0, sizeof p);
        p.row_ptr = cp.row_ptr; p.ids = cp.ids; p.cts = cp.cts;
        p.order = cp.order + c.lo; p.ndocs = (int)nd; p.counter = ctx->counters + 1 + ci;
        p.Bt = ctx->Bt; p.mw = ctx->mw; p.alpha = ctx->alpha; p.alpha_max = ctx->alpha_max;
        p.gamma = cp.gamma_dst; p.phi_ss = ctx->phi; p.docterm = cp.docterm; p.iters = cp.iters;
        p.K = K; p.KP = KP; p.ST = ST; p.max_iter = max_iter; p.tol = tol;
        p.W = W; p.nmax = nmax; p.group_bytes = gl.bytes; p.off_groups = G * gl.bytes;
        {
            const char* ce = getenv("PYLDA_COMPACT");
            p.compact = !(ce && !strcmp(ce, "0")) && !ctx->force_full;
            p.revived = ctx->counters + 14;
            p.park_nc = (c.kind == 1) ? pc.nc : 0;
            p.park_long = (c.kind == 0) ? pc.long_rows : 0;
            p.park_rec = cp.park_rec; p.park_gam = cp.park_gam; p.park_lists = cp.park_lists;
            p.park_counts = ctx->park_ctr; p.park_cap = (int)cp.D;
        }
        p.off_gam = gl.off_gam; p.off_spart = gl.off_spart; p.off_red = gl.off_red; p.off_cnt = gl.off_cnt;
        p.off_mwr = gl.off_mwr; p.off_rid = gl.off_rid; p.off_tile = gl.off_tile;
        void* args[] = {&p};
        timer.begin(ctx->stream, "%s<%d,%d> W=%d G=%d docs=%lld nmax=%d nmin=%d smem=%d grid=%lld", c.kind ? "rt" : "v2",
                    c.LK, c.J, W, G, nd, ns[c.lo], ns[c.hi - 1], smem, grid);
        CK(cudaLaunchKernel(c.fn, dim3((unsigned)grid), dim3(threads), args, (size_t)smem, ctx->stream));
        timer.end(ctx->stream);
        st->n_launches++;
        st->n_estep_launches++;
        return 0;
    };
    // Order of the launches.  Kernels that hand documents over to the narrow stages run first, then the narrow
    // stages, and the kernels that never hand over (the shared-memory classes and, when all its documents are longer
    // than 192 terms, the streaming kernel) run LAST: at that point gamma is final for every other document, so when
    // the caller's gamma buffer is page-locked its D x K copy back to the host starts here, on a second stream, and
    // crosses PCIe beside the long-document kernels instead of after them (800 MB = 15 ms at the headline config).
    // The rows of the late documents follow through the buffer's device alias (k_copy_rows).
    // (The late kernels hand their documents to estep_longc, which runs after them; the narrow stages have nothing
    // to do with documents of more than 192 terms.)
    const bool stream_parks = nstream > 0 && pc.nc > 0 && ns[(size_t)nstream - 1] <= 192;
    int nlate = 0;                                          // leading classes that never hand over to the narrow stages
    while (nlate < NC && cls[nlate].kind == 0) ++nlate;
    const long long late_lo = stream_parks ? nstream : 0;
    const long long late_hi = nlate ? cls[nlate - 1].hi : nlong;
    if (pc.nc > 0 && narrow_begin(ctx, st)) return 1;
    if (stream_parks && launch_long()) return 1;
    for (int ci = nlate; ci < NC; ++ci)
        if (launch_class(ci)) return 1;
    if (pc.nc > 0 && launch_narrow(ctx, cp, pc, max_iter, tol, st, timer)) return 1;
    const bool early = cp.early_host && cp.early_alias && cp.gamma_dst == cp.gamma && late_hi > late_lo && !timer.on;
    if (early) {
        CK(cudaEventRecord(ctx->ev_copy[0], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy[0], 0));
        CK(cudaMemcpyAsync(cp.early_host, cp.gamma, (size_t)D * K * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_copy[1], ctx->copy_stream));
    }
    if (!stream_parks && launch_long()) return 1;
    for (int ci = 0; ci < nlate; ++ci)
        if (launch_class(ci)) return 1;
    if (pc.nc > 0) {
        // documents of more than 192 terms that got down to 32 live topics (from either order of the kernels above)
        const long long nlong_docs = first_leq(192);
        if (pc.long_rows > 192 && nlong_docs > 0 &&
            launch_longc(ctx, cp, pc, ns[0], nlong_docs, max_iter, tol, st, timer)) return 1;
        if (narrow_end(ctx, st, timer)) return 1;
    }
    if (early) {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[1], 0));
        const long long n = late_hi - late_lo;
        const int blocks = (int)std::min<long long>((n + 7) / 8, (long long)ctx->prop.multiProcessorCount * 8);
        k_copy_rows<<<blocks, 256, 0, ctx->stream>>>(cp.order + late_lo, n, K, cp.gamma, cp.early_alias);
        CK(cudaGetLastError());
        st->n_launches++;
        st->gamma_rows_early = D - n;
        cp.early_done = true;
    }
    timer.report(ctx->stream);
    return 0;
}

int ensure_outputs(pylda_ctx* ctx, Corpus& cp) {
    const size_t need = (size_t)cp.D * ctx->K;
    if (cp.gamma_cap < need || !cp.gamma) {
        CK(dalloc(&cp.gamma, need));
        cp.gamma_cap = need;
    }
    if (!cp.docterm) CK(dalloc(&cp.docterm, (size_t)cp.D));
    if (!cp.iters) CK(dalloc(&cp.iters, (size_t)cp.D));
    if (!cp.park_rec) {
        CK(dalloc(&cp.park_rec, (size_t)cp.D * PARK_REC));
        CK(dalloc(&cp.park_gam, (size_t)cp.D * PARK_GAM));
        CK(dalloc(&cp.park_lists, (size_t)cp.D * PARK_LISTS));
    }
    return 0;
}

}  // namespace

extern "C" {

int pylda_abi_version(void) { return PYLDA_ABI_VERSION; }

const char* pylda_last_error(const pylda_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int pylda_create(pylda_ctx** out, int device) {
    pylda_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, "pylda_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, "pylda_create: no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, "pylda_create: device %d out of range (0..%d)", device, ndev - 1);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, "pylda_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, "cudaSetDevice: %s", cudaGetErrorString(e));
    ctx = new pylda_ctx();
    ctx->device = device;
    ctx->prop = prop;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        fail(nullptr, "cudaStreamCreate: %s", cudaGetErrorString(e));
        delete ctx;
        return 1;
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaMalloc((void**)&ctx->flat_part, (size_t)(prop.multiProcessorCount * 8 + 1) * sizeof(double));
    ctx->flat_dev = ctx->flat_part + (size_t)prop.multiProcessorCount * 8;
    cudaMallocHost((void**)&ctx->flat_host, sizeof(double));
    cudaEventCreateWithFlags(&ctx->ev_flat, cudaEventDisableTiming);
    for (auto& ev : ctx->ev_copy) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaMalloc((void**)&ctx->scal, 8 * sizeof(double));
    cudaMemset(ctx->scal, 0, 8 * sizeof(double));
    cudaMalloc((void**)&ctx->counters, 32 * sizeof(int));
    cudaMemset(ctx->counters, 0, 32 * sizeof(int));
    cudaMalloc((void**)&ctx->park_ctr, PARK_CTRS * sizeof(int));
    cudaMemset(ctx->park_ctr, 0, PARK_CTRS * sizeof(int));
    *out = ctx;
    return 0;
}

int pylda_destroy(pylda_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    free_corpus(ctx->corp[0]);
    free_corpus(ctx->corp[1]);
    free_model(ctx);
    cudaFree(ctx->scal); cudaFree(ctx->partial); cudaFree(ctx->counters); cudaFree(ctx->park_ctr);
    cudaFree(ctx->longc_tile); cudaFree(ctx->longc_cnt); cudaFree(ctx->longc_ids);
    cudaFree(ctx->flat_part); cudaFreeHost(ctx->flat_host);
    if (ctx->ev_flat) cudaEventDestroy(ctx->ev_flat);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_copy) if (ev) cudaEventDestroy(ev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

int pylda_device_name(pylda_ctx* ctx, char* out, int cap) {
    if (!ctx || !out || cap <= 0) return 1;
    snprintf(out, cap, "%s", ctx->prop.name);
    return 0;
}
int pylda_sm_count(pylda_ctx* ctx) { return ctx ? ctx->prop.multiProcessorCount : -1; }

int pylda_set_corpus(pylda_ctx* ctx, int slot, int64_t D, int64_t nnz, const int64_t* row_ptr, const int32_t* ids,
                     const int32_t* cts) {
    if (!ctx) return 1;
    if (slot < 0 || slot > 1) return fail(ctx, "pylda_set_corpus: slot must be 0 or 1");
    if (D < 0 || nnz < 0 || !row_ptr || (nnz > 0 && (!ids || !cts))) return fail(ctx, "pylda_set_corpus: bad arguments");
    if (D > 0x7fffffffLL) return fail(ctx, "pylda_set_corpus: D=%lld exceeds 2^31-1 documents per GPU", (long long)D);
    if (row_ptr[0] != 0 || row_ptr[D] != nnz) return fail(ctx, "pylda_set_corpus: row_ptr[0] must be 0 and row_ptr[D] == nnz");
    int max_n = 0;
    for (int64_t d = 0; d < D; ++d) {
        const int64_t n = row_ptr[d + 1] - row_ptr[d];
        if (n < 0) return fail(ctx, "pylda_set_corpus: row_ptr not monotone at document %lld", (long long)d);
        if (n > 0x3fffffff) return fail(ctx, "pylda_set_corpus: document %lld too long", (long long)d);
        max_n = std::max(max_n, (int)n);
    }
    int max_id = -1;
    for (int64_t i = 0; i < nnz; ++i) {
        if (ids[i] < 0) return fail(ctx, "pylda_set_corpus: negative term id at position %lld", (long long)i);
        if (cts[i] < 1) return fail(ctx, "pylda_set_corpus: count < 1 at position %lld", (long long)i);
        max_id = std::max(max_id, ids[i]);
    }
    CK(cudaSetDevice(ctx->device));
    Corpus& cp = ctx->corp[slot];
    free_corpus(cp);
    cp.D = D; cp.nnz = nnz; cp.max_id = max_id;
    // counting sort of documents by length, longest first
    std::vector<long long> bucket((size_t)max_n + 2, 0);
    for (int64_t d = 0; d < D; ++d) bucket[(size_t)(max_n - (row_ptr[d + 1] - row_ptr[d])) + 1]++;
    for (size_t i = 1; i < bucket.size(); ++i) bucket[i] += bucket[i - 1];
    std::vector<int> order((size_t)D);
    cp.n_sorted.resize((size_t)D);
    for (int64_t d = 0; d < D; ++d) {
        const int n = (int)(row_ptr[d + 1] - row_ptr[d]);
        const long long pos = bucket[(size_t)(max_n - n)]++;
        order[(size_t)pos] = (int)d;
        cp.n_sorted[(size_t)pos] = n;
    }
    CK(dalloc(&cp.row_ptr, (size_t)D + 1));
    CK(dalloc(&cp.ids, (size_t)nnz));
    CK(dalloc(&cp.cts, (size_t)nnz));
    CK(dalloc(&cp.order, (size_t)D));
    CK(cudaMemcpyAsync(cp.row_ptr, row_ptr, ((size_t)D + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
        CK(cudaMemcpyAsync(cp.ids, ids, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(cp.cts, cts, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (D) CK(cudaMemcpyAsync(cp.order, order.data(), (size_t)D * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int pylda_get_corpus(pylda_ctx* ctx, int slot, int64_t* row_ptr, int32_t* ids, int32_t* cts) {
    if (!ctx) return 1;
    if (slot < 0 || slot > 1) return fail(ctx, "pylda_get_corpus: slot must be 0 or 1");
    Corpus& cp = ctx->corp[slot];
    if (!cp.row_ptr) return fail(ctx, "pylda_get_corpus: slot %d is empty", slot);
    CK(cudaSetDevice(ctx->device));
    if (row_ptr) CK(cudaMemcpyAsync(row_ptr, cp.row_ptr, ((size_t)cp.D + 1) * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (ids && cp.nnz) CK(cudaMemcpyAsync(ids, cp.ids, (size_t)cp.nnz * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (cts && cp.nnz) CK(cudaMemcpyAsync(cts, cp.cts, (size_t)cp.nnz * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int set_alpha_host(pylda_ctx* ctx, const double* alpha_K) {
    const int K = ctx->K;
    ctx->alpha_host.assign(alpha_K, alpha_K + K);
    double amax = 0.0, asum = 0.0, lg = 0.0;
    for (int k = 0; k < K; ++k) {
        if (!(alpha_K[k] > 0.0) || !isfinite(alpha_K[k])) return fail(ctx, "alpha[%d]=%g must be finite and > 0", k, alpha_K[k]);
        amax = std::max(amax, alpha_K[k]);
        asum += alpha_K[k];
        lg += lgamma(alpha_K[k]);
    }
    ctx->alpha_max = amax;
    ctx->alpha_min = *std::min_element(alpha_K, alpha_K + K);
    ctx->alpha_sum = asum;
    ctx->lg_alpha = lg;
    ctx->alpha_term = lgamma(asum) - lg;     // variational_bayes.py:195, per document
    CK(cudaMemcpyAsync(ctx->alpha, alpha_K, (size_t)K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int pylda_set_model(pylda_ctx* ctx, int K, int V, const double* eta_KxV, const double* alpha_K) {
    if (!ctx) return 1;
    if (K < 1 || V < 1 || !eta_KxV || !alpha_K) return fail(ctx, "pylda_set_model: bad arguments");
    int LK, J;
    if (!pick_shape(K, &LK, &J)) return fail(ctx, "unsupported number of topics K=%d (max 1024)", K);
    CK(cudaSetDevice(ctx->device));
    if (K != ctx->K || V != ctx->V || !ctx->eta) {
        free_model(ctx);
        ctx->K = K; ctx->V = V; ctx->KP = (K + 1) & ~1;
        const size_t kv = (size_t)K * V, vkp = (size_t)V * ctx->KP;
        CK(dalloc(&ctx->eta, kv));
        CK(dalloc(&ctx->alpha, (size_t)K));
        CK(dalloc(&ctx->Elt, vkp));
        CK(dalloc(&ctx->Bt, vkp + 1024));     // + slack: unpredicated over-read past the last word's row
        CK(cudaMemsetAsync(ctx->Bt + vkp, 0, 1024 * sizeof(double), ctx->stream));
        CK(dalloc(&ctx->mw, (size_t)V));
        CK(dalloc(&ctx->phi, vkp));
        CK(dalloc(&ctx->kbuf, (size_t)4 * K));
        CK(dalloc(&ctx->alpha_ss, (size_t)K));
        CK(dalloc(&ctx->e_dead, (size_t)K));
        CK(dalloc(&ctx->wsum, (size_t)V));
        for (auto& cp : ctx->corp) cp.has_results = false;
    }
    CK(cudaMemcpyAsync(ctx->eta, eta_KxV, (size_t)K * V * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (set_alpha_host(ctx, alpha_K)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->model_set = true;
    return 0;
}

int pylda_set_alpha(pylda_ctx* ctx, const double* alpha_K) {
    if (!ctx) return 1;
    if (!ctx->model_set || !alpha_K) return fail(ctx, "pylda_set_alpha: no model on the device");
    CK(cudaSetDevice(ctx->device));
    if (set_alpha_host(ctx, alpha_K)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int estep_resident_impl(pylda_ctx* ctx, int slot, int max_iter, double tol, int heldout, int want_alpha_ss,
                               pylda_stats* stats, double* gamma_host_alias, double* gamma_host);

int pylda_estep_resident(pylda_ctx* ctx, int slot, int max_iter, double tol, int heldout, int want_alpha_ss,
                         pylda_stats* stats) {
    return estep_resident_impl(ctx, slot, max_iter, tol, heldout, want_alpha_ss, stats, nullptr, nullptr);
}

// gamma_host / gamma_host_alias: a page-locked caller buffer and its device-visible alias; when given, gamma reaches
// it without a D x K copy at the end of the call -- the kernels store straight into it (no narrow stages), or the
// copy starts as soon as the short documents are final and overlaps the long-document kernels (launch_estep)
static int estep_resident_impl(pylda_ctx* ctx, int slot, int max_iter, double tol, int heldout, int want_alpha_ss,
                               pylda_stats* stats, double* gamma_host_alias, double* gamma_host) {
    if (!ctx) return 1;
    if (slot < 0 || slot > 1) return fail(ctx, "pylda_estep: slot must be 0 or 1");
    if (!ctx->model_set) return fail(ctx, "pylda_estep: no model on the device (pylda_set_model)");
    Corpus& cp = ctx->corp[slot];
    if (!cp.row_ptr) return fail(ctx, "pylda_estep: corpus slot %d is empty (pylda_set_corpus)", slot);
    if (max_iter < 1) return fail(ctx, "pylda_estep: local_parameter_iteration must be >= 1 (got %d)", max_iter);
    if (cp.max_id >= ctx->V) return fail(ctx, "pylda_estep: corpus has term id %d but V=%d", cp.max_id, ctx->V);
    CK(cudaSetDevice(ctx->device));
    pylda_stats st;
    memset(&st, 0, sizeof st);
    const int K = ctx->K, V = ctx->V, KP = ctx->KP;
    if (ensure_outputs(ctx, cp)) return 1;
    // Zero-copy gamma (kernels storing straight into the caller's page-locked buffer) pays off while every kernel
    // writes whole gamma rows.  With the hand-over to the narrow stages a document's row is written twice and the
    // second time as scattered 8-byte stores, which cross PCIe one by one (measured: +25 ms per E-step on one GPU,
    // +136 ms with eight ranks sharing a host): then gamma stays on the device and leaves by one DMA copy.
    const bool zero_copy = gamma_host_alias && !want_alpha_ss && park_config(ctx).nc == 0;
    cp.gamma_dst = zero_copy ? gamma_host_alias : cp.gamma;
    cp.gamma_on_device = (cp.gamma_dst == cp.gamma);
    {
        const char* ec = getenv("PYLDA_EARLY_COPY");
        const bool on = gamma_host_alias && gamma_host && !zero_copy && !(ec && !strcmp(ec, "0"));
        cp.early_host = on ? gamma_host : nullptr;
        cp.early_alias = on ? gamma_host_alias : nullptr;
        cp.early_done = false;
    }
    const int nred = ctx->prop.multiProcessorCount * 2;
    const int nass = ctx->prop.multiProcessorCount * 2;
    if (ensure_partial(ctx, std::max((size_t)nred * NTERMS, std::max((size_t)nass * K, (size_t)2 * 64 * K)))) return 1;

    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (prepare_tables(ctx, heldout != 0, &st.n_launches)) return 1;
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    {
        const char* kv = getenv("PYLDA_KERNEL");
        const char* pv = getenv("PYLDA_PRECISION");
        const int rc = (pv && *pv && strcmp(pv, "f64")) ? launch_estep_sweep(ctx, cp, pv, max_iter, tol, &st)
                                                        : launch_estep(ctx, cp, max_iter, tol, &st);
        (void)kv;
        if (rc) return 1;
    }
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    k_reduce_terms<<<nred, 256, 0, ctx->stream>>>(ctx->phi, ctx->Elt, ctx->kbuf + 2 * K, K, V, KP, cp.docterm, cp.iters,
                                                  cp.row_ptr, cp.D, max_iter, heldout, ctx->partial);
    k_reduce_final<<<1, 256, 0, ctx->stream>>>(ctx->partial, nred, NTERMS, ctx->scal);
    st.n_launches += 2;
    ctx->have_alpha_ss = false;
    if (want_alpha_ss) {
        const int warps = 8;
        if ((size_t)warps * K * sizeof(double) > 48 * 1024)
            CK(cudaFuncSetAttribute(k_alpha_ss, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * K * (int)sizeof(double)));
        k_alpha_ss<<<nass, warps * 32, (size_t)warps * K * sizeof(double), ctx->stream>>>(cp.gamma, cp.D, K, ctx->partial);
        k_reduce_final<<<1, 256, 0, ctx->stream>>>(ctx->partial, nass, K, ctx->alpha_ss);
        st.n_launches += 2;
        ctx->have_alpha_ss = true;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (ctx->comm) {
        // [scal5] = local D so that doc_ll can add D_total * alpha_term
        const double dloc = (double)cp.D;
        CK(cudaMemcpyAsync(ctx->scal + 7, &dloc, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        // One all-reduce per E-step.  The V x K statistics go in V-range chunks of at most 1 GB inside
        // one NCCL group (4 GB at V = 1M, K = 500): NCCL pipelines them over NVLink; there is nothing to overlap
        // them with -- any document can touch any word, so the statistics are complete only when the last
        // per-document kernel has finished, and the ELBO reduction above must read the LOCAL statistics first.
        int rc = g_nccl.GroupStart();
        const size_t total = (size_t)V * KP, chunk = (size_t)1 << 27;
        for (size_t o = 0; o < total && !rc; o += chunk)
            rc = g_nccl.AllReduce(ctx->phi + o, ctx->phi + o, std::min(chunk, total - o), kNcclFloat64, kNcclSum, ctx->comm,
                                  ctx->stream);
        if (!rc) rc = g_nccl.AllReduce(ctx->scal, ctx->scal, 8, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
        if (!rc && want_alpha_ss)
            rc = g_nccl.AllReduce(ctx->alpha_ss, ctx->alpha_ss, (size_t)K, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
        const int rc2 = g_nccl.GroupEnd();
        if (rc || rc2) return fail(ctx, "NCCL all-reduce failed: %s", g_nccl.GetErrorString(rc ? rc : rc2));
    }
    ctx->phi_KV_valid = false;
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    CK(cudaMemcpyAsync(ctx->last_scal, ctx->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->comm) ctx->last_scal[7] = (double)cp.D;
    cp.has_results = true;
    cp.results_K = K;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); st.prep_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); st.kernel_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); st.post_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]); st.total_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); st.allreduce_ms = ctx->comm ? ms : 0.0;
    st.n_docs = cp.D;
    st.nnz = cp.nnz;
    st.inner_iters = (int64_t)llround(ctx->last_scal[3]);
    st.docs_at_cap = (int64_t)llround(ctx->last_scal[4]);
    st.row_trips = ctx->last_scal[5];
    {
        int rv = 0, pk[16] = {0};
        if (ctx->comm) {      // every rank must take the same decision about the full-width redo below
            const int rc = g_nccl.AllReduce(ctx->counters + 14, ctx->counters + 14, 1, /*ncclInt32*/ 2, kNcclSum, ctx->comm, ctx->stream);
            if (rc) return fail(ctx, "NCCL all-reduce failed: %s", g_nccl.GetErrorString(rc));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        cudaMemcpy(&rv, ctx->counters + 14, sizeof(int), cudaMemcpyDeviceToHost);
        if (getenv("PYLDA_TEST_FORCE_REDO") && !ctx->force_full) rv += 1;     // test hook: exercises the redo path
        cudaMemcpy(pk, ctx->park_ctr, sizeof pk, cudaMemcpyDeviceToHost);
        st.revived_docs = rv;
        st.docs_narrow_wide = (long long)pk[0] + pk[1] + pk[2] + pk[7] + pk[8];
        st.docs_narrow = (long long)pk[3] + pk[4] + pk[5] + pk[6];
        st.docs_long_compact = pk[9];
    }
    st.algo_read_bytes = 8.0 * cp.D + 8.0 * cp.nnz + 8.0 * (double)cp.nnz * K;
    st.algo_total_bytes = st.algo_read_bytes + 8.0 * (double)cp.D * K + 8.0 * (double)cp.nnz * K;
    ctx->phi_slot = slot;
    ctx->phi_heldout = heldout != 0;
    if (st.revived_docs > 0 && !ctx->force_full) {
        // The elimination of dead topics (gamma_k == alpha_k) assumes they stay dead; every kernel checks it per
        // document.  A violation has never been observed, but if one is, the results above are not the
        // reference's: redo the whole E-step with every trip at full width (a collective decision under NCCL:
        // the counter is local, so the ranks agree on it first).
        const int64_t seen = st.revived_docs;
        ctx->force_full = true;
        const int rc = estep_resident_impl(ctx, slot, max_iter, tol, heldout, want_alpha_ss, &st, gamma_host_alias, gamma_host);
        ctx->force_full = false;
        if (rc) return 1;
        st.revived_docs = seen;
    }
    if (stats) *stats = st;
    return 0;
}

int pylda_get_results(pylda_ctx* ctx, int slot, double* gamma_DxK, double* phi_ss_KxV, double* alpha_ss_K, double* doc_ll,
                      double* words_ll, int32_t* iters_D) {
    if (!ctx) return 1;
    if (slot < 0 || slot > 1) return fail(ctx, "pylda_get_results: slot must be 0 or 1");
    Corpus& cp = ctx->corp[slot];
    if (!cp.has_results) return fail(ctx, "pylda_get_results: no E-step results for slot %d", slot);
    CK(cudaSetDevice(ctx->device));
    const int K = ctx->K, V = ctx->V, KP = ctx->KP;
    if (gamma_DxK && cp.D) {
        if (!cp.gamma_on_device)
            return fail(ctx, "pylda_get_results: gamma of the last E-step was written directly to the caller's "
                             "page-locked buffer and is not on the device");
        CK(cudaMemcpyAsync(gamma_DxK, cp.gamma, (size_t)cp.D * K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (phi_ss_KxV) {
        if (!ctx->phi_KV) CK(dalloc(&ctx->phi_KV, (size_t)K * V));      // (K, V) copy: only when the caller asks for phi_ss
        if (!ctx->phi_KV_valid) {
            dim3 tb(32, 8), tg((V + 31) / 32, (K + 31) / 32);
            k_transpose_VK_to_KV<<<tg, tb, 0, ctx->stream>>>(ctx->phi, K, V, KP, ctx->phi_KV);
            CK(cudaGetLastError());
            ctx->phi_KV_valid = true;
        }
        CK(cudaMemcpyAsync(phi_ss_KxV, ctx->phi_KV, (size_t)K * V * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (alpha_ss_K) {
        if (!ctx->have_alpha_ss) return fail(ctx, "pylda_get_results: alpha_ss was not requested in the E-step call");
        CK(cudaMemcpyAsync(alpha_ss_K, ctx->alpha_ss, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (iters_D && cp.D) CK(cudaMemcpyAsync(iters_D, cp.iters, (size_t)cp.D * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const double* s = ctx->last_scal;
    if (doc_ll) *doc_ll = s[7] * ctx->alpha_term + s[0] - s[1];     // variational_bayes.py:195-199
    if (words_ll) *words_ll = s[2];                                  // :204
    return 0;
}

int pylda_estep(pylda_ctx* ctx, int slot, int K, int V, const double* eta_KxV, const double* alpha_K, int max_iter,
                double tol, int heldout, double* gamma_DxK, double* phi_ss_KxV, double* alpha_ss_K, double* doc_ll,
                double* words_ll, pylda_stats* stats) {
    if (!ctx) return 1;
    if (pylda_set_model(ctx, K, V, eta_KxV, alpha_K)) return 1;
    // page-locked (pylda_host_register / cudaHostAlloc) gamma buffer: the kernels write into it, or its copy starts early
    double* alias = nullptr;
    if (gamma_DxK) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, gamma_DxK) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            alias = (double*)at.devicePointer;
        else
            cudaGetLastError();
    }
    if (estep_resident_impl(ctx, slot, max_iter, tol, heldout, alpha_ss_K != nullptr, stats, alias, gamma_DxK)) return 1;
    const bool direct = alias && (!ctx->corp[slot].gamma_on_device || ctx->corp[slot].early_done);
    return pylda_get_results(ctx, slot, direct ? nullptr : gamma_DxK, phi_ss_KxV, alpha_ss_K, doc_ll, words_ll, nullptr);
}

int pylda_mstep_resident(pylda_ctx* ctx, double alpha_beta, double* topic_ll, double* eta_out_KxV) {
    if (!ctx) return 1;
    if (!ctx->model_set) return fail(ctx, "pylda_mstep_resident: no model on the device");
    if (!ctx->corp[0].has_results) return fail(ctx, "pylda_mstep_resident: run the training E-step first");
    if (ctx->phi_slot != 0 || ctx->phi_heldout)
        return fail(ctx, "pylda_mstep_resident: the statistics on the device come from a held-out E-step (slot %d); "
                         "run the training E-step (slot 0) again first", ctx->phi_slot);
    if (!(alpha_beta > 0.0)) return fail(ctx, "pylda_mstep_resident: alpha_beta must be > 0");
    CK(cudaSetDevice(ctx->device));
    const int K = ctx->K, V = ctx->V, KP = ctx->KP;
    // straight from the (V, KP) statistics (no (K, V) copy of them, no transpose); scratch: 2 K doubles
    double* rowterm = ctx->kbuf + 3 * K;
    {
        const int kt = (K + 31) / 32;
        const int chunks = std::max(1, std::min((V + 31) / 32, (8 * ctx->prop.multiProcessorCount + kt - 1) / kt));
        if (ensure_partial(ctx, (size_t)2 * K * chunks)) return 1;
        k_mstep_tiled<<<dim3(chunks, kt), dim3(32, 8), 0, ctx->stream>>>(ctx->eta, ctx->phi, K, V, KP, alpha_beta, ctx->partial);
        k_mstep_final<<<(K + 127) / 128, 128, 0, ctx->stream>>>(ctx->partial, K, chunks, rowterm);
    }
    CK(cudaGetLastError());
    std::vector<double> rt((size_t)K);
    CK(cudaMemcpyAsync(rt.data(), rowterm, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (eta_out_KxV)
        CK(cudaMemcpyAsync(eta_out_KxV, ctx->eta, (size_t)K * V * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // variational_bayes.py:222-224: K*(lgamma(sum alpha_beta) - sum lgamma(alpha_beta)) + sum_k rowterm_k
    double t = (double)K * (lgamma((double)V * alpha_beta) - (double)V * lgamma(alpha_beta));
    for (int k = 0; k < K; ++k) t += rt[(size_t)k];
    if (topic_ll) *topic_ll = t;
    return 0;
}

int pylda_get_eta(pylda_ctx* ctx, double* eta_KxV) {
    if (!ctx) return 1;
    if (!ctx->model_set || !eta_KxV) return fail(ctx, "pylda_get_eta: no model on the device");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(eta_KxV, ctx->eta, (size_t)ctx->K * ctx->V * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int pylda_top_words(pylda_ctx* ctx, int top, int32_t* idx_KxT, double* prob_KxT) {
    if (!ctx) return 1;
    if (!ctx->model_set) return fail(ctx, "pylda_top_words: no model on the device (pylda_set_model)");
    if (top < 1 || top > ctx->V || !idx_KxT || !prob_KxT) return fail(ctx, "pylda_top_words: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int K = ctx->K, V = ctx->V, KP = ctx->KP;
    // E_log_eta of the CURRENT eta and its per-topic logsumexp (the statistics accumulator is left alone)
    if (rowsum_psi(ctx, ctx->eta, K, V, ctx->kbuf, ctx->kbuf + K)) return 1;
    dim3 tb(32, 8), tg((V + 31) / 32, (K + 31) / 32);
    k_elog_transpose<<<tg, tb, 0, ctx->stream>>>(ctx->eta, ctx->kbuf, K, V, KP, ctx->Elt);
    const int nchunk = 64;
    if (ensure_partial(ctx, (size_t)2 * nchunk * K)) return 1;
    dim3 lg((K + 31) / 32, nchunk);
    k_lse_partial<<<lg, 256, 0, ctx->stream>>>(ctx->Elt, K, V, KP, ctx->partial, ctx->partial + (size_t)nchunk * K);
    k_lse_final<<<(K + 127) / 128, 128, 0, ctx->stream>>>(ctx->partial, ctx->partial + (size_t)nchunk * K, K, nchunk,
                                                         ctx->kbuf + 2 * K);
    CK(cudaGetLastError());
    std::string err;
    if (device_top_words(ctx->Elt, ctx->kbuf + 2 * K, K, V, KP, top, idx_KxT, prob_KxT, ctx->stream, &err))
        return fail(ctx, "pylda_top_words: %s", err.c_str());
    return 0;
}

int pylda_dirichlet_expectation(pylda_ctx* ctx, int K, int V, const double* eta_KxV, double* out_KxV) {
    if (!ctx) return 1;
    if (K < 1 || V < 1 || !eta_KxV || !out_KxV) return fail(ctx, "pylda_dirichlet_expectation: bad arguments");
    CK(cudaSetDevice(ctx->device));
    double *eta = nullptr, *ps = nullptr, *elt = nullptr, *out = nullptr;
    const int KP = (K + 1) & ~1;
    CK(dalloc(&eta, (size_t)K * V));
    CK(dalloc(&ps, (size_t)2 * K));
    CK(dalloc(&elt, (size_t)V * KP));
    CK(dalloc(&out, (size_t)K * V));
    CK(cudaMemcpyAsync(eta, eta_KxV, (size_t)K * V * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (rowsum_psi(ctx, eta, K, V, ps, nullptr)) return 1;
    dim3 tb(32, 8), tg((V + 31) / 32, (K + 31) / 32);
    k_elog_transpose<<<tg, tb, 0, ctx->stream>>>(eta, ps, K, V, KP, elt);
    k_transpose_VK_to_KV<<<tg, tb, 0, ctx->stream>>>(elt, K, V, KP, out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_KxV, out, (size_t)K * V * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(eta); cudaFree(ps); cudaFree(elt); cudaFree(out);
    return 0;
}

int pylda_special(pylda_ctx* ctx, int which, int64_t n, const double* x, double* out) {
    if (!ctx) return 1;
    if (which < 0 || which > 4 || n < 0 || (n > 0 && (!x || !out))) return fail(ctx, "pylda_special: bad arguments");
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    double *dx = nullptr, *dy = nullptr;
    CK(dalloc(&dx, (size_t)n));
    CK(dalloc(&dy, (size_t)n));
    CK(cudaMemcpyAsync(dx, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_special<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(which, n, dx, dy);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dy, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(dx); cudaFree(dy);
    return 0;
}

int pylda_host_register(pylda_ctx* ctx, void* ptr, int64_t bytes) {
    if (!ctx) return 1;
    if (!ptr || bytes <= 0) return fail(ctx, "pylda_host_register: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    return 0;
}
int pylda_host_unregister(pylda_ctx* ctx, void* ptr) {
    if (!ctx) return 1;
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostUnregister(ptr));
    return 0;
}

int pylda_comm_allreduce_sum(pylda_ctx* ctx, double* buf, int64_t n) {
    if (!ctx) return 1;
    if (n < 0 || (n > 0 && !buf)) return fail(ctx, "pylda_comm_allreduce_sum: bad arguments");
    if (!ctx->comm || n == 0) return 0;              // single rank: the sum over ranks is the input
    CK(cudaSetDevice(ctx->device));
    double* d = nullptr;
    CK(dalloc(&d, (size_t)n));
    CK(cudaMemcpyAsync(d, buf, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int rc = g_nccl.AllReduce(d, d, (size_t)n, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
    if (rc) { cudaFree(d); return fail(ctx, "ncclAllReduce: %s", g_nccl.GetErrorString(rc)); }
    CK(cudaMemcpyAsync(buf, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    return 0;
}

int pylda_comm_unique_id(char id_out[PYLDA_NCCL_ID_BYTES]) {
    std::string err;
    if (!load_nccl(&err)) return fail(nullptr, "%s", err.c_str());
    ncclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc) return fail(nullptr, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
    memcpy(id_out, id.internal, PYLDA_NCCL_ID_BYTES);
    return 0;
}

int pylda_comm_init(pylda_ctx* ctx, int n_ranks, int rank, const char id_in[PYLDA_NCCL_ID_BYTES]) {
    if (!ctx) return 1;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !id_in) return fail(ctx, "pylda_comm_init: bad arguments");
    std::string err;
    if (!load_nccl(&err)) return fail(ctx, "%s", err.c_str());
    CK(cudaSetDevice(ctx->device));
    if (ctx->comm) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; }
    ncclUniqueId id;
    memcpy(id.internal, id_in, PYLDA_NCCL_ID_BYTES);
    const int rc = g_nccl.CommInitRank(&ctx->comm, n_ranks, id, rank);
    if (rc) { ctx->comm = nullptr; return fail(ctx, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc)); }
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    return 0;
}

}  // extern "C"
