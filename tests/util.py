import os

import numpy

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["ap200_k10", "ap200_k10_warm3", "nips24_k200", "syn96_k50", "zipf48_k100"]


def load_golden(name):
    from pylda_b200 import synthetic
    z = numpy.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    K, V = int(g["K"]), int(g["V"])
    g["K"], g["V"] = K, V
    if "eta" not in g:
        g["eta"] = synthetic.initial_eta(K, V, int(g["eta_seed"]))
    phi = numpy.zeros((K, V))
    phi[:, g["phi_cols"]] = g["phi_ss_cols"]
    g["phi_ss"] = phi
    g["doc_ll"] = float(g["doc_ll"])
    return g


def max_rel(a, b, floor=0.0):
    a = numpy.asarray(a, dtype=numpy.float64)
    b = numpy.asarray(b, dtype=numpy.float64)
    den = numpy.maximum(numpy.abs(b), floor) if floor > 0 else numpy.abs(b)
    with numpy.errstate(divide="ignore", invalid="ignore"):
        r = numpy.where(den > 0, numpy.abs(a - b) / den, numpy.abs(a - b))
    return float(r.max()) if r.size else 0.0


# north_star tolerance: gamma / beta statistics / ELBO within 1e-5 relative (fp64)
RTOL = 1e-5
# phi_ss entries span 300 orders of magnitude; entries below this absolute floor are compared
# absolutely (an entry of 1e-200 carries no information at 1e-5 of the ELBO)
PHI_FLOOR = 1e-12
