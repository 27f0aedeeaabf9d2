"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy, fp64) of PyLDA's VB E-step.

This module is the *checker*.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it.  The product
path (pylda_b200/) never does: it fails loudly when the CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so
this restatement is pinned against the reference's own code executed in the
build container through oracle/ref_shim.py (tests/test_oracle_pin.py) and against
the fixtures that oracle/make_golden.py generated from that code
(tests/golden/*.npz).

Every function cites the reference lines it follows (paths relative to
/root/reference).  The arithmetic is deliberately the reference's formulation
(log-space phi, scipy logsumexp, one numpy call per reference call) so that its
CPU cost profile is the reference's too -- it doubles as the "port" CPU baseline.
"""
import numpy
import scipy.special


def compute_dirichlet_expectation(dirichlet_parameter):
    """inferencer.py:15-18 -- psi(x) - psi(sum x) (1-D) or row-wise (2-D)."""
    if dirichlet_parameter.ndim == 1:
        return scipy.special.psi(dirichlet_parameter) - scipy.special.psi(numpy.sum(dirichlet_parameter))
    return scipy.special.psi(dirichlet_parameter) - scipy.special.psi(numpy.sum(dirichlet_parameter, 1))[:, numpy.newaxis]


def csr_from_parsed(word_ids, word_cts):
    """Pack the reference's parsed-corpus type (variational_bayes.py:98-130:
    word_ids[d] int (n_d,), word_cts[d] int (1,n_d)) into CSR.  Integer work: bit-exact."""
    D = len(word_ids)
    row_ptr = numpy.zeros(D + 1, dtype=numpy.int64)
    for d in range(D):
        row_ptr[d + 1] = row_ptr[d] + len(word_ids[d])
    ids = numpy.zeros(int(row_ptr[-1]), dtype=numpy.int32)
    cts = numpy.zeros(int(row_ptr[-1]), dtype=numpy.int32)
    for d in range(D):
        ids[row_ptr[d]:row_ptr[d + 1]] = word_ids[d]
        cts[row_ptr[d]:row_ptr[d + 1]] = numpy.asarray(word_cts[d]).reshape(-1)
    return row_ptr, ids, cts


def parsed_from_csr(row_ptr, ids, cts):
    """Inverse of csr_from_parsed -- the reference's (word_ids, word_cts) lists."""
    word_ids, word_cts = [], []
    for d in range(len(row_ptr) - 1):
        a, b = int(row_ptr[d]), int(row_ptr[d + 1])
        word_ids.append(numpy.asarray(ids[a:b], dtype=numpy.int64))
        word_cts.append(numpy.asarray(cts[a:b], dtype=numpy.int64)[numpy.newaxis, :])
    return word_ids, word_cts


def e_step(row_ptr, ids, cts, eta, alpha, max_iter=50, tol=1e-6, heldout=False,
           doc_order=None, return_iters=False):
    """variational_bayes.py:132-216 on a CSR corpus.

    eta (K,V) f64, alpha (K,) f64.  Returns a dict with
      gamma (D,K), phi_ss (K,V), doc_ll (document_log_likelihood, :195-199),
      words_ll (:204; 0 when heldout is False), iters (D,) inner-iteration counts.
    doc_order replaces numpy.random.permutation(D) (:159); it only changes fp
    summation order of the accumulators.
    """
    K, V = eta.shape
    D = len(row_ptr) - 1
    document_log_likelihood = 0.0
    words_log_likelihood = 0.0
    phi_sufficient_statistics = numpy.zeros((K, V))                       # :147
    gamma_values = numpy.zeros((D, K)) + alpha[numpy.newaxis, :] + 1.0 * V / K   # :150
    E_log_eta = compute_dirichlet_expectation(eta)                         # :152
    if heldout:                                                            # :154-155
        E_log_prob_eta = E_log_eta - scipy.special.logsumexp(E_log_eta, axis=1)[:, numpy.newaxis]
    iters = numpy.zeros(D, dtype=numpy.int32)
    if doc_order is None:
        doc_order = range(D)
    alpha_term = scipy.special.gammaln(numpy.sum(alpha)) - numpy.sum(scipy.special.gammaln(alpha))
    for doc_id in doc_order:
        a, b = int(row_ptr[doc_id]), int(row_ptr[doc_id + 1])
        term_ids = ids[a:b]
        term_counts = cts[a:b].astype(numpy.int64)[numpy.newaxis, :]       # (1, n_d) as in :121
        total_word_count = numpy.sum(term_counts)                          # :162
        gamma_values[doc_id, :] = alpha + 1.0 * total_word_count / K       # :165
        n_d = term_ids.shape[0]
        log_counts_col = numpy.log(term_counts.transpose())
        for gamma_iteration in range(max_iter):                            # :174
            log_phi = E_log_eta[:, term_ids].T + numpy.tile(scipy.special.psi(gamma_values[[doc_id], :]), (n_d, 1))  # :177
            log_phi -= scipy.special.logsumexp(log_phi, axis=1)[:, numpy.newaxis]   # :182
            gamma_update = alpha + numpy.array(numpy.sum(numpy.exp(log_phi + log_counts_col), axis=0))  # :185
            mean_change = numpy.mean(abs(gamma_update - gamma_values[doc_id, :]))   # :187
            gamma_values[doc_id, :] = gamma_update                         # :188
            iters[doc_id] = gamma_iteration + 1
            if mean_change <= tol:                                         # :189-190
                break
        document_log_likelihood += alpha_term                              # :195
        document_log_likelihood += numpy.sum(scipy.special.gammaln(gamma_values[doc_id, :])) - scipy.special.gammaln(numpy.sum(gamma_values[doc_id, :]))  # :197
        document_log_likelihood -= numpy.sum(numpy.dot(term_counts, numpy.exp(log_phi) * log_phi))  # :199
        if heldout:                                                        # :202-204
            words_log_likelihood += numpy.sum(numpy.exp(log_phi.T + numpy.log(term_counts)) * E_log_prob_eta[:, term_ids])
        phi_sufficient_statistics[:, term_ids] += numpy.exp(log_phi + log_counts_col).T   # :207
    out = dict(gamma=gamma_values, phi_ss=phi_sufficient_statistics,
               doc_ll=float(document_log_likelihood), words_ll=float(words_log_likelihood))
    if return_iters:
        out["iters"] = iters
    return out


def m_step(eta, gamma, phi_ss, alpha_beta):
    """variational_bayes.py:218-235.  Topic ELBO terms are taken from the *old* eta
    (:222-224) before eta <- phi_ss + alpha_beta (:226).  Returns
    (topic_log_likelihood, new_eta, alpha_sufficient_statistics)."""
    K = eta.shape[0]
    topic_log_likelihood = K * (scipy.special.gammaln(numpy.sum(alpha_beta)) - numpy.sum(scipy.special.gammaln(alpha_beta)))
    topic_log_likelihood += numpy.sum(numpy.sum(scipy.special.gammaln(eta), axis=1) - scipy.special.gammaln(numpy.sum(eta, axis=1)))
    new_eta = phi_ss + alpha_beta
    alpha_ss = scipy.special.psi(gamma) - scipy.special.psi(numpy.sum(gamma, axis=1)[:, numpy.newaxis])
    alpha_ss = numpy.sum(alpha_ss, axis=0)
    return float(topic_log_likelihood), new_eta, alpha_ss


def optimize_hyperparameters(alpha, alpha_ss, number_of_documents, hyper_parameter_iteration=100,
                             hyper_parameter_decay_factor=0.9, hyper_parameter_maximum_decay=10,
                             hyper_parameter_converge_threshold=1e-6):
    """variational_bayes.py:277-324 -- linear-time Newton step on asymmetric alpha.
    Keeps the reference's quirk: `sum_1_h = 1.0 / alpha_hessian` (:292) is a *vector*
    (no numpy.sum), hence `c` (:295) is a vector too."""
    alpha = numpy.array(alpha, dtype=numpy.float64)
    alpha_update = alpha
    decay = 0
    for alpha_iteration in range(hyper_parameter_iteration):
        alpha_sum = numpy.sum(alpha)
        alpha_gradient = number_of_documents * (scipy.special.psi(alpha_sum) - scipy.special.psi(alpha)) + alpha_ss
        alpha_hessian = -number_of_documents * scipy.special.polygamma(1, alpha)
        sum_g_h = numpy.sum(alpha_gradient / alpha_hessian)
        sum_1_h = 1.0 / alpha_hessian
        z = number_of_documents * scipy.special.polygamma(1, alpha_sum)
        c = sum_g_h / (1.0 / z + sum_1_h)
        while True:
            singular_hessian = False
            step_size = numpy.power(hyper_parameter_decay_factor, decay) * (alpha_gradient - c) / alpha_hessian
            if numpy.any(alpha <= step_size):
                singular_hessian = True
            else:
                alpha_update = alpha - step_size
            if singular_hessian:
                decay += 1
                if decay > hyper_parameter_maximum_decay:
                    break
            else:
                break
        mean_change = numpy.mean(abs(alpha_update - alpha))
        alpha = alpha_update
        if mean_change <= hyper_parameter_converge_threshold:
            break
    return alpha


def learning_trace(row_ptr, ids, cts, eta0, alpha0, alpha_beta, iterations, max_iter=50, tol=1e-6):
    """variational_bayes.py:239-261 repeated `iterations` times (launch_train.py:196-197):
    E-step, M-step, alpha update every iteration (interval 1, :59).  Returns the list of
    joint ELBOs plus the final (eta, alpha, gamma)."""
    eta, alpha = numpy.array(eta0), numpy.array(alpha0)
    D = len(row_ptr) - 1
    trace = []
    gamma = None
    for _ in range(iterations):
        r = e_step(row_ptr, ids, cts, eta, alpha, max_iter, tol)
        gamma = r["gamma"]
        topic_ll, eta, alpha_ss = m_step(eta, gamma, r["phi_ss"], alpha_beta)
        alpha = optimize_hyperparameters(alpha, alpha_ss, D)
        trace.append(r["doc_ll"] + topic_ll)
    return trace, eta, alpha, gamma
