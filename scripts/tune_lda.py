#!/usr/bin/env python
"""Tuning aid for the warm state: EM iterations 1..4 on a corpus drawn from an LDA model (bench.py's warm_lda),
then per-class timings of the E-step of EM iteration 5 and -- on a sample of documents, in numpy on the host --
how many topics are still alive (gamma_k != alpha_k) trip by trip under that model.
   python scripts/tune_lda.py [docs] [sample]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, scipy.special
from pylda_b200 import native, synthetic
from pylda_b200.variational_bayes import VariationalBayes
import bench

D = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
K, V = 100, 100000
row_ptr, ids, cts = bench.load_corpus(D, V, 4321, kind="lda")
alpha = numpy.full(K, 1.0 / K)
ctx = native.EStepContext(0)
ctx.set_corpus(0, row_ptr, ids, cts)
shell = VariationalBayes()
shell._number_of_topics = K
shell._number_of_documents = D
shell._alpha_alpha = alpha.copy()
ctx.set_model(synthetic.initial_eta(K, V, 0), alpha)
for em in range(4):
    st = ctx.estep_resident(0, 50, 1e-6, want_alpha_ss=True)
    print("EM %d: kernel %.2f ms trips %.2f narrow %d/%d" % (em + 1, st["kernel_ms"], st["inner_iters"] / D, st["docs_narrow_wide"], st["docs_narrow"]), flush=True)
    alpha_ss = ctx.get_results(0, gamma=False, phi=False, alpha_ss=True)["alpha_ss"]
    ctx.mstep_resident(1.0 / V, want_eta=False)
    shell.optimize_hyperparameters(alpha_ss)
    ctx.set_alpha(shell._alpha_alpha)
a5 = shell._alpha_alpha.copy()
print("alpha after 4 EM iterations: min %.4g max %.4g sum %.4g" % (a5.min(), a5.max(), a5.sum()))
for _ in range(2):
    ctx.estep_resident(0, 50, 1e-6)
os.environ["PYLDA_PROFILE_CLASSES"] = "1"
st = ctx.estep_resident(0, 50, 1e-6)
os.environ.pop("PYLDA_PROFILE_CLASSES", None)
print("EM 5: kernel %.2f ms trips %.2f at cap %d narrow %d/%d long-compact %d revived %d" % (
    st["kernel_ms"], st["inner_iters"] / D, st["docs_at_cap"], st["docs_narrow_wide"], st["docs_narrow"], st["docs_long_compact"], st["revived_docs"]), flush=True)
it = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
nd = numpy.diff(row_ptr)
edges = [0, 24, 48, 96, 192, 272, 4096]
for a, b in zip(edges[:-1], edges[1:]):
    m = (nd > a) & (nd <= b)
    if m.any():
        print("  n in (%d,%d]: docs %d mean trips %.1f p10 %d p50 %d p90 %d" % ((a, b, m.sum(), it[m].mean()) + tuple(numpy.percentile(it[m], [10, 50, 90]).astype(int))))
# live topics trip by trip, numpy, on a sample
eta = ctx.get_eta()
Elog = scipy.special.psi(eta) - scipy.special.psi(eta.sum(1))[:, None]
rs = numpy.random.RandomState(0)
pick = rs.choice(D, S, replace=False)
T = 50
live = -numpy.ones((S, T), dtype=numpy.int32)
for i, d in enumerate(pick):
    a, b = row_ptr[d], row_ptr[d + 1]
    El = Elog[:, ids[a:b]].T
    B = numpy.exp(El - El.max(1)[:, None])
    c = cts[a:b].astype(float)
    g = a5 + c.sum() / K
    for t in range(T):
        e = numpy.exp(scipy.special.psi(g))
        gn = a5 + e * ((c / (B @ e)) @ B)
        ch = numpy.mean(abs(gn - g))
        g = gn
        live[i, t] = (g != a5).sum()
        if ch <= 1e-6:
            break
for a, b in zip(edges[:-1], edges[1:]):
    m = (nd[pick] > a) & (nd[pick] <= b)
    if not m.any():
        continue
    print("n in (%d,%d]: sample %d" % (a, b, m.sum()))
    for t in [0, 2, 4, 6, 8, 10, 12, 15, 20, 25, 30, 40, 49]:
        x = live[m, t]
        run = x >= 0
        if run.any():
            x = x[run]
            print("   trip %2d: running %.3f live mean %.1f p50 %d p90 %d" % (t + 1, run.mean(), x.mean(), numpy.percentile(x, 50), numpy.percentile(x, 90)))
