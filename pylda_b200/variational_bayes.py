"""Drop-in VariationalBayes (mirror of /root/reference/variational_bayes.py:55-356) whose
e_step runs on a B200 through the C ABI in include/pylda_b200.h.

What stays host numpy, exactly as in the reference: parse_data (:98-130), m_step (:218-235),
optimize_hyperparameters (:277-324, including the vector-valued `sum_1_h` quirk at :292),
export_beta / export_gamma (:326-356).  What moved to the device: everything inside e_step
(:132-216) including compute_dirichlet_expectation (:152).  There is no CPU fallback: e_step
raises if the CUDA library or a B200 is missing.
"""
import os
import sys
import time

import numpy
import scipy.special

from .inferencer import Inferencer, compute_dirichlet_expectation
from . import corpus_cache, distributed, native


def pack_parsed_corpus(parsed_corpus):
    """(word_ids, word_cts) lists of variational_bayes.py:98-130 -> CSR (row_ptr int64,
    ids int32, cts int32).  Integer work: the token indexing is preserved bit-exactly."""
    word_ids, word_cts = parsed_corpus
    assert len(word_ids) == len(word_cts)
    lengths = numpy.fromiter((len(w) for w in word_ids), dtype=numpy.int64, count=len(word_ids))
    row_ptr = numpy.zeros(len(word_ids) + 1, dtype=numpy.int64)
    numpy.cumsum(lengths, out=row_ptr[1:])
    if len(word_ids):
        ids = numpy.concatenate([numpy.asarray(w).reshape(-1) for w in word_ids]).astype(numpy.int32)
        cts = numpy.concatenate([numpy.asarray(c).reshape(-1) for c in word_cts]).astype(numpy.int32)
    else:
        ids = numpy.zeros(0, dtype=numpy.int32)
        cts = numpy.zeros(0, dtype=numpy.int32)
    assert ids.shape[0] == row_ptr[-1] == cts.shape[0]
    return row_ptr, ids, cts


class VariationalBayes(Inferencer):
    def __init__(self, hyper_parameter_optimize_interval=1):
        Inferencer.__init__(self, hyper_parameter_optimize_interval)   # :58-67
        self._native = None
        self._train_uploaded = False
        self._alpha_ss_device = None

    # ---- self._eta / self._gamma: plain attributes for every caller, but after a resident EM
    # iteration (learning() below) the current values live in HBM and are copied back only when
    # somebody reads them (export_*, pickling, a direct e_step call) ----
    @property
    def _eta(self):
        d = self.__dict__
        if d.get("_eta_stale") and d.get("_native") is not None:
            d["_eta_host"] = d["_native"].get_eta()
            d["_eta_stale"] = False
        return d.get("_eta_host")

    @_eta.setter
    def _eta(self, value):
        d = self.__dict__
        d["_eta_host"] = value
        d["_eta_stale"] = False
        d["_model_on_device"] = False          # the device copy (if any) is no longer the model

    @property
    def _gamma(self):
        d = self.__dict__
        if d.get("_gamma_stale") and d.get("_native") is not None:
            rows = d["_native"].get_results(0, gamma=True, phi=False)["gamma"]
            lo, hi = d.get("_gamma_rows", (0, rows.shape[0]))
            # multi-process: a collective -- every rank must read _gamma at the same point (materialize())
            d["_gamma_host"] = self._gather_rows(rows, lo, hi, self._number_of_documents)
            d["_gamma_stale"] = False
        return d.get("_gamma_host")

    @_gamma.setter
    def _gamma(self, value):
        self.__dict__["_gamma_host"] = value
        self.__dict__["_gamma_stale"] = False

    # ---- the object must stay picklable (launch_train.py:203-204): drop the device handle ----
    def __getstate__(self):
        self._eta, self._gamma = self._eta, self._gamma        # bring the current values to the host
        state = dict(self.__dict__)
        state["_native"] = None
        state["_train_uploaded"] = False
        state["_alpha_ss_device"] = None
        state["_model_on_device"] = False
        state.pop("_last_parsed_csr", None)                    # (attribute of older versions)
        state.pop("_train_shard", None)
        if isinstance(state.get("_parsed_corpus"), corpus_cache.LazyParsed):
            state["_parsed_corpus"] = tuple(state["_parsed_corpus"]._materialize())
        return state

    def _context(self):
        """The device context of this process.  Launched as one process per GPU (RANK / WORLD_SIZE /
        LOCAL_RANK in the environment, e.g. by torchrun) the ranks are joined with NCCL: every rank
        parses the same corpus, runs the E-step on its own nnz-balanced shard of documents, and the
        library all-reduces the K x V statistics, the ELBO scalars and the alpha statistics.
        The device keeps the gamma rows of the local shard (self._gamma_rows = (lo, hi)); self._gamma gathers
        all D rows over NCCL when it is read."""
        if self._native is None:
            rank, size, local = distributed.world()
            self._native = native.EStepContext(local)
            self._rank, self._world = rank, size
            if size > 1:
                if os.environ.get("PYTHONHASHSEED", "random") == "random":
                    raise RuntimeError("pylda_b200: multi-process runs need the same PYTHONHASHSEED on every rank "
                                       "(type ids come from set() iteration order, inferencer.py:63-65)")
                uid = distributed.exchange_unique_id(rank, size, native.EStepContext.comm_unique_id)
                self._native.comm_init(size, rank, uid)
            self._train_uploaded = False
        return self._native

    def _upload_train(self, ctx):
        if not self._train_uploaded:
            if getattr(self, "_train_shard", None) is not None:
                lo, hi, shard = self._train_shard              # read from the on-disk CSR cache: this rank's rows only
            else:
                if getattr(self, "_train_csr", None) is None:
                    self._train_csr = pack_parsed_corpus(self._parsed_corpus)
                lo, hi, shard = self._shard(self._train_csr)
            ctx.set_corpus(0, *shard)
            self._gamma_rows = (lo, hi)
            self._train_uploaded = True

    def _shard(self, csr):
        """(lo, hi, csr shard) of this rank; the whole corpus when single-process."""
        row_ptr, ids, cts = csr
        D = len(row_ptr) - 1
        if getattr(self, "_world", 1) <= 1:
            return 0, D, csr
        b = native.shard_bounds(row_ptr, self._world)
        lo, hi = int(b[self._rank]), int(b[self._rank + 1])
        return lo, hi, native.shard_csr(row_ptr, ids, cts, lo, hi)

    def _initialize(self, corpus, vocab, number_of_topics, alpha_alpha, alpha_beta):
        # :82-95
        Inferencer._initialize(self, vocab, number_of_topics, alpha_alpha, alpha_beta)
        self.__dict__.pop("_train_shard", None)
        cache_dir = os.environ.get("PYLDA_CSR_CACHE")
        if cache_dir:
            # on-disk CSR (corpus_cache.py): rank 0 parses once, every rank maps the files and reads its shard only
            csr = None
            self._parsed_corpus = self._parse_cached(corpus, cache_dir)
            self._number_of_documents = self._parsed_corpus.number_of_documents
        else:
            self._parsed_corpus, csr = self._parse(corpus)
            self._number_of_documents = len(self._parsed_corpus[0])
        self._gamma = numpy.zeros((self._number_of_documents, self._number_of_topics)) \
            + self._alpha_alpha[numpy.newaxis, :] + 1.0 * self._number_of_types / self._number_of_topics
        # the only random draw that affects VB results (:95); same global-RNG call as the reference
        self._eta = numpy.random.gamma(100., 1. / 100., (self._number_of_topics, self._number_of_types))
        # (the native parser already produced the CSR of exactly this corpus; checked, never trusted blindly)
        if csr is not None and len(csr[0]) - 1 != self._number_of_documents:
            csr = None
        if cache_dir:
            self._train_csr = None
        else:
            self._train_csr = csr if csr is not None else pack_parsed_corpus(self._parsed_corpus)
        self._train_uploaded = False
        rank, size, _ = distributed.world()
        if size > 1:
            # every rank drew its own eta0; rank 0's draw becomes the model of all ranks
            if rank != 0:
                self._eta[:] = 0.0
            self._context().allreduce_sum(self._eta)

    def _parse_cached(self, corpus, cache_dir):
        """The training corpus through the on-disk CSR cache: parsed by rank 0 only when the entry is missing; this
        rank keeps its own shard (self._train_shard) and a lazy stand-in for the (word_ids, word_cts) lists."""
        rank, size, _ = distributed.world()
        key, vocab_sha1, corpus_sha1 = corpus_cache.cache_key(corpus, self._index_to_type)
        entry = corpus_cache.open_entry(cache_dir, key)
        if entry is None:
            if rank == 0:
                parsed, csr = self._parse(corpus)
                dropped = len(corpus) - len(parsed[0])
                corpus_cache.save(cache_dir, key, csr if csr is not None else pack_parsed_corpus(parsed), dropped,
                                  vocab_sha1, corpus_sha1)
            else:
                corpus_cache.wait_for(cache_dir, key)
            entry = corpus_cache.open_entry(cache_dir, key)
        else:
            print("successfully parse %d documents..." % entry[0]["D"])
        self._train_shard = corpus_cache.load_shard(entry, rank, size)
        return corpus_cache.LazyParsed(entry)

    def parse_data(self, corpus):
        # :98-130 -- per document: unique in-vocabulary type ids (first-seen order) and counts;
        # documents with no in-vocabulary token are dropped with a warning.
        return self._parse(corpus)[0]

    def _parse(self, corpus):
        """parse_data plus the CSR form of the same corpus when the native parser produced it (else None).
        Nothing is kept on the object: a CSR left over from an earlier call can never be mistaken for
        the corpus of a later one.
        Fast path: the native parser of the C ABI (pylda_parse_corpus, multi-threaded host code, same
        semantics bit for bit); the loop below remains for non-ASCII text and as its specification."""
        if os.environ.get("PYLDA_NATIVE_PARSE", "1") != "0" and isinstance(corpus, (list, tuple)):
            parsed = native.parse_corpus(corpus, self._index_to_type)
            if parsed is not None:
                row_ptr, ids, cts, dropped = parsed
                for _ in range(dropped):
                    sys.stderr.write("warning: document collapsed during parsing")
                ids64, cts64 = ids.astype(numpy.int64), cts.astype(numpy.int64)
                bounds = row_ptr[1:-1]
                word_ids = numpy.split(ids64, bounds) if len(row_ptr) > 1 else []
                word_cts = [c[numpy.newaxis, :] for c in numpy.split(cts64, bounds)] if len(row_ptr) > 1 else []
                for done in range(10000, len(word_ids) + 1, 10000):
                    print("successfully parse %d documents..." % done)
                print("successfully parse %d documents..." % len(word_ids))
                return (word_ids, word_cts), (row_ptr, ids, cts)
        doc_count = 0
        word_ids, word_cts = [], []
        lookup = self._type_to_index
        for document_line in corpus:
            counts = {}
            for token in document_line.split():
                type_id = lookup.get(token)
                if type_id is None:
                    continue
                counts[type_id] = counts.get(type_id, 0) + 1
            if not counts:
                sys.stderr.write("warning: document collapsed during parsing")
                continue
            word_ids.append(numpy.array(list(counts.keys())))
            word_cts.append(numpy.array(list(counts.values()))[numpy.newaxis, :])
            doc_count += 1
            if doc_count % 10000 == 0:
                print("successfully parse %d documents..." % doc_count)
        assert len(word_ids) == len(word_cts)
        print("successfully parse %d documents..." % doc_count)
        return (word_ids, word_cts), None

    def e_step(self, parsed_corpus=None, local_parameter_iteration=50, local_parameter_converge_threshold=1e-6):
        """:132-216.  Train branch (parsed_corpus is None): sets self._gamma and returns
        (document_log_likelihood, phi_sufficient_statistics (K,V)).  Held-out branch: returns
        (words_log_likelihood, gamma_values) and leaves self._gamma untouched."""
        return self._e_step_impl(parsed_corpus, local_parameter_iteration, local_parameter_converge_threshold, None)

    def _e_step_impl(self, parsed_corpus, local_parameter_iteration, local_parameter_converge_threshold, csr):
        ctx = self._context()
        heldout = parsed_corpus is not None
        multi = self._world > 1
        if heldout:
            word_ids, word_cts = parsed_corpus[0], parsed_corpus[1]
            assert len(word_ids) == len(word_cts)
            if csr is None or len(csr[0]) - 1 != len(word_ids):
                csr = pack_parsed_corpus((word_ids, word_cts))
            lo, hi, shard = self._shard(csr)
            ctx.set_corpus(1, *shard)
            slot, number_of_documents = 1, len(word_ids)
        else:
            self._upload_train(ctx)
            lo, hi = self._gamma_rows
            slot, number_of_documents = 0, self._number_of_documents
        # the reference visits documents in numpy.random.permutation order (:159); the order only
        # changes fp summation order, but the draw keeps the global RNG stream in step with it
        numpy.random.permutation(number_of_documents)
        out = ctx.estep(slot, self._eta, self._alpha_alpha, local_parameter_iteration,
                        local_parameter_converge_threshold, heldout=heldout,
                        want_gamma=True, want_phi=not heldout, want_alpha_ss=multi and not heldout)
        self._last_estep_stats = out["stats"]
        self.__dict__["_model_on_device"] = False      # pylda_estep re-uploads eta/alpha from the host
        gamma = self._gather_rows(out["gamma"], lo, hi, number_of_documents)
        if not heldout:
            self._gamma = gamma                        # all D documents, as in the reference (:212)
            self._alpha_ss_device = out["alpha_ss"]    # summed over ranks by the library (None single-process)
            return out["doc_ll"], out["phi_ss"]
        return out["words_ll"], gamma

    def _gather_rows(self, rows, lo, hi, number_of_documents):
        """Multi-process: every rank holds the gamma rows [lo, hi) of its document shard; the reference's
        callers (export_gamma :343-356, launch_test.py:95-97, the pickled model) expect all D rows.  The
        shards are summed into a zero-padded D x K buffer over NCCL (a collective: every rank calls it)."""
        if getattr(self, "_world", 1) <= 1:
            return rows
        full = numpy.zeros((number_of_documents, rows.shape[1]), dtype=numpy.float64)
        full[lo:hi, :] = rows
        return self._context().allreduce_sum(full)

    def m_step(self, phi_sufficient_statistics):
        # :218-235 -- topic terms from the OLD eta, then eta <- phi_ss + alpha_beta
        topic_log_likelihood = self._number_of_topics * (
            scipy.special.gammaln(numpy.sum(self._alpha_beta)) - numpy.sum(scipy.special.gammaln(self._alpha_beta)))
        topic_log_likelihood += numpy.sum(numpy.sum(scipy.special.gammaln(self._eta), axis=1)
                                          - scipy.special.gammaln(numpy.sum(self._eta, axis=1)))
        self._eta = phi_sufficient_statistics + self._alpha_beta
        assert self._eta.shape == (self._number_of_topics, self._number_of_types)
        if getattr(self, "_alpha_ss_device", None) is not None:
            # multi-process: self._gamma is only this rank's shard; the same statistic (:232-233) was
            # computed on every rank's device from its gamma rows and all-reduced by the library
            alpha_sufficient_statistics = self._alpha_ss_device
        else:
            alpha_sufficient_statistics = scipy.special.psi(self._gamma) \
                - scipy.special.psi(numpy.sum(self._gamma, axis=1)[:, numpy.newaxis])
            alpha_sufficient_statistics = numpy.sum(alpha_sufficient_statistics, axis=0)
        return topic_log_likelihood, alpha_sufficient_statistics

    def _resident_ok(self):
        """Resident EM is used when nothing in the subclass changes the E/M-step and the topic prior is
        the reference's uniform vector (inferencer.py:58); PYLDA_RESIDENT=0 forces the host M-step."""
        return (os.environ.get("PYLDA_RESIDENT", "1") != "0"
                and type(self).e_step is VariationalBayes.e_step and type(self).m_step is VariationalBayes.m_step
                and numpy.all(self._alpha_beta == self._alpha_beta[0]))

    def _learning_resident(self):
        """learning() with the model resident in HBM: device E-step, device
        M-step (:222-226), alpha statistics (:232-233) from the device, the reference's Newton update of
        alpha on the host (K numbers).  No K x V or D x K array crosses PCIe; eta and gamma are fetched
        lazily when read.  Same arithmetic, same printed line."""
        self._counter += 1
        clock_e_step = time.time()
        ctx = self._context()
        self._upload_train(ctx)
        if not self.__dict__.get("_model_on_device"):
            ctx.set_model(self._eta, self._alpha_alpha)
        numpy.random.permutation(self._number_of_documents)            # RNG stream of :159
        self._last_estep_stats = ctx.estep_resident(0, 50, 1e-6, want_alpha_ss=True)
        res = ctx.get_results(0, gamma=False, phi=False, alpha_ss=True)
        clock_e_step = time.time() - clock_e_step
        clock_m_step = time.time()
        topic_log_likelihood, _ = ctx.mstep_resident(float(self._alpha_beta[0]), want_eta=False)
        d = self.__dict__
        d["_eta_stale"] = d["_gamma_stale"] = True
        d["_model_on_device"] = True
        self._alpha_ss_device = res["alpha_ss"] if self._world > 1 else None
        if self._hyper_parameter_optimize_interval > 0 and self._counter % self._hyper_parameter_optimize_interval == 0:
            self.optimize_hyperparameters(res["alpha_ss"])
            ctx.set_alpha(self._alpha_alpha)
        clock_m_step = time.time() - clock_m_step
        joint_log_likelihood = res["doc_ll"] + topic_log_likelihood
        print("e_step and m_step of iteration %d finished in %d and %d seconds respectively with log likelihood %g"
              % (self._counter, clock_e_step, clock_m_step, joint_log_likelihood))
        return joint_log_likelihood

    def learning(self):
        # :239-261
        if self._resident_ok():
            return self._learning_resident()
        self._counter += 1
        clock_e_step = time.time()
        document_log_likelihood, phi_sufficient_statistics = self.e_step()
        clock_e_step = time.time() - clock_e_step
        clock_m_step = time.time()
        topic_log_likelihood, alpha_sufficient_statistics = self.m_step(phi_sufficient_statistics)
        if self._hyper_parameter_optimize_interval > 0 and self._counter % self._hyper_parameter_optimize_interval == 0:
            self.optimize_hyperparameters(alpha_sufficient_statistics)
        clock_m_step = time.time() - clock_m_step
        joint_log_likelihood = document_log_likelihood + topic_log_likelihood
        print("e_step and m_step of iteration %d finished in %d and %d seconds respectively with log likelihood %g"
              % (self._counter, clock_e_step, clock_m_step, joint_log_likelihood))
        return joint_log_likelihood

    def inference(self, corpus):
        # :263-271
        parsed_corpus, csr = self._parse(corpus)
        words_log_likelihood, corpus_gamma_values = self._e_step_impl(parsed_corpus, 50, 1e-6, csr)
        return words_log_likelihood, corpus_gamma_values

    def materialize(self):
        """Bring eta and gamma (all D rows) to the host.  After resident EM iterations they live in HBM and
        are fetched when read; in a multi-process run fetching gamma is a collective, so every rank calls
        this before rank 0 exports or pickles the model (launch_train does)."""
        return self._eta, self._gamma

    def optimize_hyperparameters(self, alpha_sufficient_statistics, hyper_parameter_iteration=100,
                                 hyper_parameter_decay_factor=0.9, hyper_parameter_maximum_decay=10,
                                 hyper_parameter_converge_threshold=1e-6):
        # :277-324 -- Newton update of the asymmetric alpha (Minka's linear-time form) with step
        # decay.  NOTE the reference's quirk, kept on purpose so ELBO traces match: `sum_1_h` is
        # 1/hessian element-wise, NOT summed (:292), so `c` is a K-vector (:295).
        assert alpha_sufficient_statistics.shape == (self._number_of_topics,)
        D = self._number_of_documents
        alpha_update = self._alpha_alpha
        decay = 0
        for _ in range(hyper_parameter_iteration):
            alpha_sum = numpy.sum(self._alpha_alpha)
            gradient = D * (scipy.special.psi(alpha_sum) - scipy.special.psi(self._alpha_alpha)) \
                + alpha_sufficient_statistics
            hessian = -D * scipy.special.polygamma(1, self._alpha_alpha)
            if numpy.any(numpy.isinf(gradient)) or numpy.any(numpy.isnan(gradient)):
                print("illegal alpha gradient vector", gradient)
            sum_g_h = numpy.sum(gradient / hessian)
            sum_1_h = 1.0 / hessian
            z = D * scipy.special.polygamma(1, alpha_sum)
            c = sum_g_h / (1.0 / z + sum_1_h)
            while True:
                step_size = numpy.power(hyper_parameter_decay_factor, decay) * (gradient - c) / hessian
                assert self._alpha_alpha.shape == step_size.shape
                if numpy.any(self._alpha_alpha <= step_size):
                    decay += 1
                    if decay > hyper_parameter_maximum_decay:
                        break
                else:
                    alpha_update = self._alpha_alpha - step_size
                    break
            mean_change = numpy.mean(abs(alpha_update - self._alpha_alpha))
            self._alpha_alpha = alpha_update
            if mean_change <= hyper_parameter_converge_threshold:
                break
        return

    def export_beta(self, exp_beta_path, top_display=-1):
        # :326-341
        if self.__dict__.get("_native") is not None and os.environ.get("PYLDA_DEVICE_EXPORT", "1") != "0":
            # device path: per-topic sort on the device, only the displayed words come back
            ctx = self._native
            if not self.__dict__.get("_model_on_device"):
                ctx.set_model(self._eta, self._alpha_alpha)
                self.__dict__["_model_on_device"] = True
            top = self._number_of_types if top_display <= 0 else min(top_display, self._number_of_types)
            idx, prob = ctx.top_words(top)
            with open(exp_beta_path, 'w') as output:
                for topic_index in range(self._number_of_topics):
                    output.write("==========\t%d\t==========\n" % topic_index)
                    for type_index, p in zip(idx[topic_index], prob[topic_index]):
                        output.write("%s\t%g\n" % (self._index_to_type[int(type_index)], p))
            return
        E_log_eta = compute_dirichlet_expectation(self._eta)
        with open(exp_beta_path, 'w') as output:
            for topic_index in range(self._number_of_topics):
                output.write("==========\t%d\t==========\n" % topic_index)
                row = E_log_eta[topic_index, :]
                beta_probability = numpy.exp(row - scipy.special.logsumexp(row))
                for rank, type_index in enumerate(reversed(numpy.argsort(beta_probability)), 1):
                    output.write("%s\t%g\n" % (self._index_to_type[type_index], beta_probability[type_index]))
                    if top_display > 0 and rank >= top_display:
                        break

    def export_gamma(self, exp_gamma_path, top_display=-1):
        # :343-356
        exp_gamma = self._gamma / numpy.sum(self._gamma, axis=1)[:, numpy.newaxis]
        with open(exp_gamma_path, 'w') as output:
            for document_index in range(exp_gamma.shape[0]):
                fields = []
                for rank, topic_index in enumerate(reversed(numpy.argsort(exp_gamma[document_index, :])), 1):
                    fields.append("%d:%g" % (topic_index, exp_gamma[document_index, topic_index]))
                    if top_display > 0 and rank >= top_display:
                        break
                output.write("%s\n" % "\t".join(fields))


if __name__ == "__main__":
    print("not implemented...")
