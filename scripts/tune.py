#!/usr/bin/env python
"""Tuning aid: one cold E-step (+ one repeat) on the headline corpus shape with per-class timings.
   PYLDA_CLASSES=... python scripts/tune.py [docs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
from pylda_b200 import native, synthetic
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
MAXIT = int(os.environ.get("TUNE_MAXITER", "50"))
D = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
K, V = 100, 100000
row_ptr, ids, cts = bench.load_corpus(D, V, 1236)
ctx = native.EStepContext(0)
ctx.set_corpus(0, row_ptr, ids, cts)
ctx.set_model(synthetic.initial_eta(K, V, 0), numpy.full(K, 1.0 / K))
os.environ.pop("PYLDA_PROFILE_CLASSES", None)
for _ in range(2):
    ctx.estep_resident(0, MAXIT, 1e-6)
os.environ["PYLDA_PROFILE_CLASSES"] = "1"
st = ctx.estep_resident(0, MAXIT, 1e-6)
os.environ.pop("PYLDA_PROFILE_CLASSES", None)
ks = [ctx.estep_resident(0, MAXIT, 1e-6)["kernel_ms"] for _ in range(3)]
r = ctx.get_results(0, gamma=False, phi=False)
print("classes=%s D=%d kernel_ms=%.3f (min of 3) trips=%.2f doc_ll=%.10e narrow %d/%d long-compact %d revived %d" % (
    os.environ.get("PYLDA_CLASSES", "default"), D, min(ks), st["inner_iters"] / D, r["doc_ll"], st["docs_narrow_wide"], st["docs_narrow"], st["docs_long_compact"], st["revived_docs"]), flush=True)
