"""The premise of the kernels' dead-topic elimination, checked on the CPU with a
numpy model of the product-form recurrence: once gamma_k == alpha_k bit for bit the topic stays dead for
the rest of the document's trips, and dropping dead topics from the norms / column sums changes gamma by
rounding only (and the trip count not at all).  The GPU implementation is checked against the oracle and
against its own full-width path in tests/test_estep_gpu.py."""
import numpy
import scipy.special as sp

from tests.util import load_golden


def _trips(B, c, alpha, eliminate, max_iter=50, tol=1e-6):
    K = alpha.shape[0]
    gam = alpha + c.sum() / K
    live = numpy.ones(K, dtype=bool)
    e = numpy.exp(sp.psi(gam))
    revivals = 0
    it = 0
    while True:
        if eliminate:
            norm = B[:, live] @ e[live]
            s = numpy.zeros(K)
            s[live] = B[:, live].T @ (c / norm)
        else:
            norm = B @ e
            s = B.T @ (c / norm)
        gn = numpy.where(live, alpha + e * s, alpha) if eliminate else alpha + e * s
        dead_now = gn == alpha
        if not eliminate:
            revivals += int((~live & ~dead_now).sum())          # a topic that was bitwise dead moved again
            live = ~dead_now
        else:
            live = live & ~dead_now
        it += 1
        change = numpy.mean(numpy.abs(gn - gam))
        gam = gn
        if change <= tol or it >= max_iter:
            return gam, it, revivals, int(live.sum())
        e = numpy.exp(sp.psi(gam))


def _check_fixture(name, n_docs):
    g = load_golden(name)
    eta, alpha = g["eta"], g["alpha"]
    Elog = sp.psi(eta) - sp.psi(eta.sum(axis=1))[:, None]
    total_revivals = 0
    died = 0
    for d in range(min(n_docs, len(g["row_ptr"]) - 1)):
        a, b = int(g["row_ptr"][d]), int(g["row_ptr"][d + 1])
        if a == b:
            continue
        El = Elog[:, g["ids"][a:b]].T
        B = numpy.exp(El - El.max(axis=1)[:, None])
        c = g["cts"][a:b].astype(numpy.float64)
        full, it_full, revivals, live_end = _trips(B, c, alpha, eliminate=False)
        fast, it_fast, _, _ = _trips(B, c, alpha, eliminate=True)
        total_revivals += revivals
        died += alpha.shape[0] - live_end
        assert it_full == it_fast
        assert numpy.max(numpy.abs(fast - full) / full) <= 1e-12
        # and the model itself is the reference's recurrence (fixture gamma comes from the reference)
        assert numpy.max(numpy.abs(full - g["gamma"][d]) / g["gamma"][d]) <= 1e-9
    return total_revivals, died


def test_elimination_premise_on_zipf_k100():
    revivals, died = _check_fixture("zipf48_k100", 48)
    assert revivals == 0 and died > 1000          # most of the 48 x 100 topics end up bitwise dead


def test_elimination_premise_on_synthetic_k50_warm():
    revivals, died = _check_fixture("syn96_k50", 60)
    assert revivals == 0


def test_elimination_is_inert_where_no_topic_dies():
    revivals, died = _check_fixture("ap200_k10_warm3", 60)      # alpha ~ 0.1: exp(psi(alpha)) is not negligible
    assert revivals == 0 and died == 0
