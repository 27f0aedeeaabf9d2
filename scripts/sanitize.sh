#!/bin/bash
# compute-sanitizer memcheck + racecheck on small E-steps that cover every kernel (register tile, narrow stages,
# shared-memory tile, streaming with hand-over, compact stage for long documents, hybrid cluster kernel), the device M-step and the top-words sort
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys, numpy
sys.path.insert(0, os.getcwd())
from pylda_b200 import native, synthetic
V = 3000
a = synthetic.synthetic_corpus(60, V, seed=99, length="zipf")
b = synthetic.synthetic_corpus(2, V, seed=98, length="poisson", mean_len=3000)
row_ptr = numpy.concatenate([a[0], a[0][-1] + b[0][1:]])
ids, cts = numpy.concatenate([a[1], b[1]]), numpy.concatenate([a[2], b[2]])
ctx = native.EStepContext(0)
ctx.set_corpus(0, row_ptr, ids, cts)
for K in (100, 10, 50, 200, 500):
    eta = synthetic.initial_eta(K, V, 1)
    alpha = numpy.full(K, 1.0 / K)
    for kern in ("default", "v2", "hybrid"):
        if kern == "default":
            os.environ.pop("PYLDA_KERNEL", None)
        else:
            os.environ["PYLDA_KERNEL"] = kern
        out = ctx.estep(0, eta, alpha, 30, 1e-6, heldout=True, want_alpha_ss=True)
        st = out["stats"]
        print("K=%d" % K, kern, out["doc_ll"], st["n_estep_launches"], "narrow", st["docs_narrow_wide"], st["docs_narrow"], flush=True)
    os.environ.pop("PYLDA_KERNEL", None)
    ctx.estep_resident(0, 30, 1e-6, want_alpha_ss=True)          # resident EM iteration: device M-step, top words
    ctx.mstep_resident(1.0 / V, want_eta=False)
    ctx.top_words(5)
    st = ctx.estep_resident(0, 50, 1e-6)                         # the model after one M-step: long documents reach estep_longc
    print("K=%d after the M-step: long-compact %d narrow %d" % (K, st["docs_long_compact"], st["docs_narrow"]), flush=True)
ctx.close()
PY
for tool in memcheck racecheck; do
  echo "=== $tool" | tee -a gpurun_out/sanitize.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py 2>&1 | grep -v "^=========     at\|^=========     by\|^=========         in" | tail -40 | tee -a gpurun_out/sanitize.log
done
