"""Design evidence (CPU, numpy): how many topics of a document are still alive (gamma_k != alpha_k)
after each trip of the fixed point, on a sample of the headline corpus (config 3, cold state).
Drives the width schedule of the E-step kernels.  Not on the product path."""
import sys, numpy, scipy.special
sys.path.insert(0, '.')
from pylda_b200.synthetic import synthetic_corpus, initial_eta

D = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
K, V = 100, 100000
row_ptr, ids, cts = synthetic_corpus(D, V, 1237, length="zipf")
eta = initial_eta(K, V)
Elog = scipy.special.psi(eta) - scipy.special.psi(eta.sum(1))[:, None]
alpha = numpy.full(K, 1.0 / K)
nd = numpy.diff(row_ptr)
print("docs", D, "nnz", len(ids), "mean n", nd.mean(), "max n", nd.max())
edges = [0, 24, 48, 96, 192, 272, 512, 1024, 4096]
print("class share of docs / rows:")
for a, b in zip(edges[:-1], edges[1:]):
    m = (nd > a) & (nd <= b)
    print("  n in (%d,%d]: docs %.4f rows %.4f" % (a, b, m.mean(), nd[m].sum() / nd.sum()))
T = 50
live = numpy.zeros((D, T), dtype=numpy.int32)
iters = numpy.zeros(D, dtype=numpy.int32)
for d in range(D):
    a, b = row_ptr[d], row_ptr[d + 1]
    El = Elog[:, ids[a:b]].T
    B = numpy.exp(El - El.max(1)[:, None])
    c = cts[a:b].astype(float)
    g = alpha + c.sum() / K
    for t in range(T):
        e = numpy.exp(scipy.special.psi(g))
        norm = B @ e
        gn = alpha + e * ((c / norm) @ B)
        ch = numpy.mean(abs(gn - g))
        g = gn
        live[d, t] = (g != alpha).sum()
        iters[d] = t + 1
        if ch <= 1e-6:
            live[d, t + 1:] = -1
            break
numpy.savez("/tmp/sim_live.npz", live=live, iters=iters, nd=nd)
print("mean trips", iters.mean(), "at cap", (iters == 50).mean())
for a, b in zip(edges[:-1], edges[1:]):
    m = (nd > a) & (nd <= b)
    if not m.any():
        continue
    L = live[m]
    print("n in (%d,%d]: docs %d mean trips %.1f" % (a, b, m.sum(), iters[m].mean()))
    for t in [0, 1, 2, 3, 4, 5, 6, 8, 10, 15, 20, 30, 40, 49]:
        x = L[:, t]
        x = x[x >= 0]
        if len(x):
            print("   trip %2d: running %.3f live mean %.1f p50 %d p90 %d max %d" % (t + 1, len(x) / m.sum(), x.mean(), numpy.percentile(x, 50), numpy.percentile(x, 90), x.max()))
# work model: sum over doc-trips of n * width(live), for a few width schedules
def work(widths):
    tot = 0
    for d in range(D):
        w = 100
        for t in range(iters[d]):
            tot += nd[d] * w                       # trip t+1 runs at the width decided after trip t
            l = live[d, t]
            for ww in widths:
                if l <= ww:
                    w = min(w, ww)
    return tot
full = float((nd * iters).sum() * 100)
for sched in [[], [32], [64, 32], [64, 32, 16], [64, 32, 16, 8], [64, 48, 32, 24, 16, 12, 8]]:
    print("width schedule", sched, "work fraction %.3f" % (work(sched) / full))
