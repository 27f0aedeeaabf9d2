/*
 * pylda_b200 -- C ABI of the B200-native variational-Bayes E-step for LDA.
 *
 * Drop-in boundary for ONE path of kzhai/PyLDA (reference, pure Python, no FFI of its own):
 *     variational_bayes.py:132-216   VariationalBayes.e_step
 *     inferencer.py:15-18            compute_dirichlet_expectation
 * The reference-side binding (a ctypes stub inside VariationalBayes.e_step) is shown in
 * INTEGRATION.md; pylda_b200/native.py is that stub, pylda_b200/variational_bayes.py the
 * class that uses it.
 *
 * Conventions
 *   - plain pointers and sizes only; every host buffer is caller-owned, C-contiguous, and
 *     only read/written during the call.  The library owns all device memory, streams,
 *     events and the NCCL communicator.
 *   - every function returns 0 on success, non-zero on error; the message is available
 *     from pylda_last_error().  Nothing is thrown across the ABI.
 *   - one context per process and GPU (one process per GPU; ranks are joined with
 *     pylda_comm_init).  Calls on one context must be serialised by the caller.
 *   - matrices use the REFERENCE's layouts: eta / phi_ss are (K, V) row-major doubles
 *     (variational_bayes.py:95,147), gamma is (D, K) row-major (variational_bayes.py:150),
 *     alpha is (K,) (inferencer.py:57).  The (K,V) <-> (V,K) transposition happens on the
 *     device inside the library.
 *   - there is no CPU fallback: without a usable CUDA device pylda_create fails.
 */
#ifndef PYLDA_B200_H
#define PYLDA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYLDA_ABI_VERSION 10
#define PYLDA_NCCL_ID_BYTES 128

typedef struct pylda_ctx pylda_ctx;

/* Filled by the E-step entry points.  n_docs, nnz, times, bytes and launch counts are LOCAL to the
 * rank; inner_iters, docs_at_cap and row_trips are summed over ranks once pylda_comm_init was called. */
typedef struct pylda_stats {
    int64_t n_docs;            /* documents processed by this context (local shard)          */
    int64_t nnz;               /* (doc, term) pairs processed                                 */
    int64_t inner_iters;       /* sum over documents of gamma iterations (:174 loop trips)    */
    int64_t docs_at_cap;       /* documents that stopped at max_iter instead of converging    */
    double  prep_ms;           /* E_log_eta producer + table build (inferencer.py:15-18)      */
    double  kernel_ms;         /* per-document E-step kernels, CUDA-event time on our stream  */
    double  post_ms;           /* ELBO reduction, transposes, (all-reduce)                    */
    double  total_ms;          /* device time of the whole call incl. H2D/D2H issued by us    */
    double  algo_read_bytes;   /* sum_d 8 + 8 n_d + 8 n_d K                     */
    double  algo_total_bytes;  /* read + sum_d 8 K + 8 n_d K                    */
    int32_t n_launches;        /* kernels launched by this call                               */
    int32_t n_estep_launches;  /* of which per-document E-step kernels                        */
    int64_t docs_resident;     /* documents whose tile was staged in shared memory            */
    int64_t docs_streamed;     /* documents whose tile was re-streamed from L2/HBM            */
    double  row_trips;         /* sum_d n_d * iters_d: 4*KP*row_trips = fp64 flops of the mat-vecs */
    int64_t revived_docs;      /* documents in which a topic eliminated as dead (gamma_k == alpha_k) would have
                                  come back (summed over ranks).  0 in every corpus seen so far; when it is not, the
                                  library has already redone the E-step at full width and the results are those */
    int64_t docs_narrow_wide;  /* documents handed to the 32- / 16-column narrow stages (at most 32 / 16 topics alive;
                                  a document that passes both is counted twice)                                  */
    int64_t docs_narrow;       /* documents that went through the 8-column narrow stage (at most 8 alive)      */
    double  allreduce_ms;      /* NCCL all-reduce of the V x K statistics, ELBO scalars and alpha statistics
                                  (CUDA-event time on our stream; part of post_ms; 0 on a single rank)          */
    int64_t gamma_rows_early;  /* rows of gamma whose copy to a page-locked caller buffer (pylda_estep) started before
                                  the long-document kernels and overlapped them; 0 = gamma left at the end       */
    int64_t docs_long_compact; /* documents of more than 192 terms finished by the compact stage for long documents
                                  (at most 32 topics alive)                                                      */
} pylda_stats;

/* ABI version of the loaded library (== PYLDA_ABI_VERSION of the header it was built from). */
int pylda_abi_version(void);

/* Create a context on CUDA device `device` (cudaSetDevice ordinal).  Fails (non-zero) when
 * no sm_100 device is usable.  *out is NULL on failure. */
int pylda_create(pylda_ctx** out, int device);
int pylda_destroy(pylda_ctx* ctx);

/* Last error message of `ctx` (or of the failed pylda_create when ctx is NULL).
 * Owned by the library; valid until the next call on the same context. */
const char* pylda_last_error(const pylda_ctx* ctx);

/* Upload a parsed corpus (the reference's (word_ids, word_cts) of variational_bayes.py:98-130
 * packed to CSR: row_ptr[D+1] int64, ids[nnz] int32 term ids unique within a document,
 * cts[nnz] int32 counts >= 1).  slot 0 = training corpus (self._parsed_corpus),
 * slot 1 = held-out corpus (the parsed_corpus argument of e_step).  Under multi-GPU every
 * rank uploads ITS OWN shard of documents.  Copies; the caller may free its arrays. */
int pylda_set_corpus(pylda_ctx* ctx, int slot, int64_t D, int64_t nnz,
                     const int64_t* row_ptr, const int32_t* ids, const int32_t* cts);

/* Read the corpus back from the device (token indexing must round-trip bit-exactly). */
int pylda_get_corpus(pylda_ctx* ctx, int slot, int64_t* row_ptr, int32_t* ids, int32_t* cts);

/* The whole reference call  VariationalBayes.e_step(parsed_corpus, local_parameter_iteration,
 * local_parameter_converge_threshold)  (variational_bayes.py:132) with HOST buffers:
 *   eta_KxV, alpha_K     inputs (self._eta, self._alpha_alpha)
 *   heldout              0: train branch (:212-214), 1: held-out branch (:154-155,:202-204,:216)
 *   gamma_DxK            out, nullable  (gamma_values)
 *   phi_ss_KxV           out, nullable  (phi_sufficient_statistics; summed over ranks)
 *   alpha_ss_K           out, nullable  (sum_d psi(gamma_dk) - psi(sum_k gamma_dk), the alpha
 *                        statistics m_step needs, variational_bayes.py:232-233; summed over ranks)
 *   doc_ll               out: document_log_likelihood (:195-199), summed over ranks
 *   words_ll             out: words_log_likelihood (:204), 0 when heldout == 0
 * Under multi-GPU (after pylda_comm_init) this is a collective: every rank calls it. */
int pylda_estep(pylda_ctx* ctx, int slot, int K, int V,
                const double* eta_KxV, const double* alpha_K,
                int max_iter, double tol, int heldout,
                double* gamma_DxK, double* phi_ss_KxV, double* alpha_ss_K,
                double* doc_ll, double* words_ll, pylda_stats* stats);

/* The same path split so that inputs can stay resident in HBM between calls
 * (bench.py `value` leg; also what a device-resident M-step feeds):
 *   pylda_set_model     H2D of eta/alpha into the context
 *   pylda_estep_resident  E_log_eta producer + E-step kernels + ELBO reduction (+ all-reduce),
 *                       results stay on the device
 *   pylda_get_results   D2H of whatever the caller wants (any pointer may be NULL) */
int pylda_set_model(pylda_ctx* ctx, int K, int V, const double* eta_KxV, const double* alpha_K);
int pylda_estep_resident(pylda_ctx* ctx, int slot, int max_iter, double tol, int heldout,
                         int want_alpha_ss, pylda_stats* stats);
int pylda_get_results(pylda_ctx* ctx, int slot, double* gamma_DxK, double* phi_ss_KxV,
                      double* alpha_ss_K, double* doc_ll, double* words_ll, int32_t* iters_D);

/* Device M-step on resident statistics (variational_bayes.py:218-226):
 * topic_ll from the OLD eta, then eta <- phi_ss + alpha_beta (scalar prior, inferencer.py:58).
 * eta stays on the device for the next pylda_estep_resident; eta_out_KxV is nullable. */
int pylda_mstep_resident(pylda_ctx* ctx, double alpha_beta, double* topic_ll, double* eta_out_KxV);
/* Copy the resident eta (K, V) back to the host (after resident EM iterations). */
int pylda_get_eta(pylda_ctx* ctx, double* eta_KxV);
/* Replace alpha on the device (after the host Newton update, variational_bayes.py:277-324). */
int pylda_set_alpha(pylda_ctx* ctx, const double* alpha_K);

/* export_beta on the device (variational_bayes.py:326-341): for every topic the `top` most probable words under
 * beta_kv = exp(E_log_eta[k,v] - logsumexp_v E_log_eta[k,:]) of the model currently on the device (pylda_set_model /
 * after pylda_mstep_resident), most probable first.  idx_KxT: word (type) ids, prob_KxT: their probabilities, both
 * K x top row-major host buffers.  Replaces K host argsorts over V entries and the K x V copy-back they need. */
int pylda_top_words(pylda_ctx* ctx, int top, int32_t* idx_KxT, double* prob_KxT);

/* compute_dirichlet_expectation (inferencer.py:15-18) for a (K,V) matrix on the device.
 * Exposed for parity tests of the device digamma. */
int pylda_dirichlet_expectation(pylda_ctx* ctx, int K, int V, const double* eta_KxV, double* out_KxV);

/* Elementwise device special functions used on the path (parity tests against scipy):
 * which = 0: digamma(x); 1: exp(digamma(x)) (shifted form, c = 0); 2: lgamma(x); 3: Newton reciprocal
 * 1/x; 4: exp(digamma(x)) as the second-generation kernel evaluates it. */
int pylda_special(pylda_ctx* ctx, int which, int64_t n, const double* x, double* out);

/* Multi-GPU: one process per GPU.  Rank 0 calls pylda_comm_unique_id, ships the bytes to the
 * other ranks by any host channel, then every rank calls pylda_comm_init.  Afterwards
 * pylda_estep / pylda_estep_resident all-reduce (NCCL, sum, f64) the K x V statistics and
 * the ELBO scalars; gamma stays sharded. */
int pylda_comm_unique_id(char id_out[PYLDA_NCCL_ID_BYTES]);
int pylda_comm_init(pylda_ctx* ctx, int n_ranks, int rank, const char id[PYLDA_NCCL_ID_BYTES]);
/* In-place sum over ranks of a host buffer of n doubles (H2D, ncclAllReduce, D2H); a no-op without a
 * communicator.  Used by the class layer to make rank 0's random eta0 draw (variational_bayes.py:95)
 * the model of every rank. */
int pylda_comm_allreduce_sum(pylda_ctx* ctx, double* buf, int64_t n);

/* Page-lock / unlock a caller-owned host buffer (cudaHostRegister, mapped) so that the H2D/D2H copies of
 * pylda_estep run at full PCIe rate.  Optional: pageable buffers work, only slower.  When the gamma_DxK
 * argument of pylda_estep is page-locked, the D x K copy of gamma does not wait for the end of the call: the
 * kernels of the long documents run last and the copy of everything else crosses PCIe beside them (the long
 * documents' rows follow through the buffer's device alias; pylda_stats.gamma_rows_early).  When the hand-over
 * to the narrow stages is off (alpha too large for topics to die, or PYLDA_PARK=0) and alpha_ss_K is NULL, the
 * kernels store gamma directly into the buffer instead. */
int pylda_host_register(pylda_ctx* ctx, void* ptr, int64_t bytes);
int pylda_host_unregister(pylda_ctx* ctx, void* ptr);

/* Host-side corpus ingestion (no device involved): the native counterpart of VariationalBayes.parse_data
 * (variational_bayes.py:98-130).  `text` holds one document per '\n'-terminated line, `vocab` one word per
 * '\n'-separated line in type-id order (self._index_to_type).  Tokens are split on ASCII whitespace exactly
 * like str.split(); out-of-vocabulary tokens are skipped (:107-108); per document the distinct type ids
 * come in first-seen order with their counts (:110-113); documents without an in-vocabulary token are
 * dropped (:115-117).  n_threads <= 0: all host threads.  The result is read back with pylda_parsed_dims /
 * pylda_parsed_copy (row_ptr[D+1] int64, ids[nnz] int32, cts[nnz] int32) and released with pylda_parsed_free. */
typedef struct pylda_parsed pylda_parsed;
int pylda_parse_corpus(const char* text, int64_t text_len, const char* vocab, int64_t vocab_len, int n_threads,
                       pylda_parsed** out);
int pylda_parsed_dims(const pylda_parsed* p, int64_t* D, int64_t* nnz, int64_t* dropped);
int pylda_parsed_copy(const pylda_parsed* p, int64_t* row_ptr, int32_t* ids, int32_t* cts);
int pylda_parsed_free(pylda_parsed* p);

/* Introspection for benches/tests. */
int pylda_device_name(pylda_ctx* ctx, char* out, int cap);
int pylda_sm_count(pylda_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PYLDA_B200_H */
