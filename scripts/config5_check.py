#!/usr/bin/env python
"""One-off functional check at the BASELINE config-5 model size (V = 1M, K = 500: every (V, K) table is
4 GB, > 2^32 bytes) with a reduced number of documents: invariants only (no oracle at this size)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
from pylda_b200 import native, synthetic
K, V, D = 500, 1000000, int(sys.argv[1]) if len(sys.argv) > 1 else 20000
t = time.time()
row_ptr, ids, cts = synthetic.synthetic_corpus(D, V, seed=1238, length="poisson", mean_len=100)
eta = numpy.random.RandomState(0).gamma(100., 0.01, (K, V))
print("inputs in %.1fs, nnz=%d, eta %.1f GB" % (time.time() - t, len(ids), eta.nbytes / 1e9), flush=True)
alpha = numpy.full(K, 1.0 / K)
ctx = native.EStepContext(0)
ctx.set_corpus(0, row_ptr, ids, cts)
t = time.time()
ctx.set_model(eta, alpha)
for kern in os.environ.get("C5_KERNELS", "").split(","):      # tuning aid: time other long-document kernels first
    if kern:
        os.environ["PYLDA_KERNEL"] = kern
        for _ in range(2):
            s2 = ctx.estep_resident(0, 50, 1e-6)
        print("PYLDA_KERNEL=%s: kernel %.1f ms, %.0f docs/s, streamed %d" % (kern, s2["kernel_ms"], D / (s2["total_ms"] * 1e-3), s2["docs_streamed"]), flush=True)
        del os.environ["PYLDA_KERNEL"]
st = ctx.estep_resident(0, 50, 1e-6, want_alpha_ss=True)
st = ctx.estep_resident(0, 50, 1e-6, want_alpha_ss=True)
print("E-step %.1f ms kernel, %.1f ms prep, %.1f ms post; wall incl. H2D %.1fs" % (st["kernel_ms"], st["prep_ms"], st["post_ms"], time.time() - t), flush=True)
res = ctx.get_results(0, gamma=True, phi=True, alpha_ss=True, iters=True)
N = numpy.add.reduceat(cts.astype(numpy.float64), row_ptr[:-1])
assert numpy.allclose(res["gamma"].sum(axis=1), alpha.sum() + N, rtol=1e-11)
assert abs(res["phi_ss"].sum() - cts.sum()) <= 1e-9 * cts.sum()
cf = numpy.bincount(ids, weights=cts.astype(numpy.float64), minlength=V)
assert numpy.allclose(res["phi_ss"].sum(axis=0), cf, rtol=1e-9, atol=1e-9)
assert numpy.isfinite(res["doc_ll"]) and res["iters"].max() <= 50
topic_ll, _ = ctx.mstep_resident(1.0 / V, want_eta=False)
assert numpy.isfinite(topic_ll)
print("config-5 model size OK: docs/s %.0f, doc_ll %.6e, topic_ll %.6e, mean trips %.1f, streamed %d resident %d" % (
    D / (st["total_ms"] * 1e-3), res["doc_ll"], topic_ll, st["inner_iters"] / D, st["docs_streamed"], st["docs_resident"]))
