"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED
reference (variational_bayes.py / inferencer.py, loaded through oracle/ref_shim.py).

Run in the build container only (needs /root/reference):

    PYTHONHASHSEED=0 python oracle/make_golden.py [--trace] [--nips-full [--only-nips-full]]

Vocabulary ids come from set() iteration order (inferencer.py:63-65), so the script
re-executes itself with PYTHONHASHSEED=0 when that is not set.  Fixtures hold the
CSR-packed parsed corpus (integer, bit-exact), eta0, alpha and the reference's
outputs, so the GPU-box tests never need the reference.
"""
import hashlib
import os
import sys
import tarfile
import tempfile

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import estep_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _extract(tar_rel, member_dir):
    tmp = tempfile.mkdtemp(prefix="pylda_ref_")
    with tarfile.open(os.path.join(ref_shim.REFERENCE_ROOT, tar_rel)) as t:
        t.extractall(tmp)
    return os.path.join(tmp, member_dir)


def _read_docs(path):
    # launch_train.py:102-107
    docs = []
    with open(path) as f:
        for line in f:
            docs.append(line.strip().lower())
    return docs


def _read_vocab(path):
    # launch_train.py:110-116
    vocab = []
    with open(path) as f:
        for line in f:
            vocab.append(line.strip().lower().split()[0])
    return list(set(vocab))


def _run_case(name, vb_mod, train_docs, vocab, K, heldout_docs=None, em_warm=0):
    """One reference e_step (train branch) [+ held-out branch] -> fixture dict."""
    V_guess = len(set(vocab))
    numpy.random.seed(0)
    lda = vb_mod.VariationalBayes()
    lda._initialize(train_docs, vocab, K, 1.0 / K, 1.0 / V_guess)    # launch_train.py:118-124,194
    for _ in range(em_warm):
        lda.learning()
    eta0 = lda._eta.copy()
    alpha0 = lda._alpha_alpha.copy()
    row_ptr, ids, cts = O.csr_from_parsed(*lda._parsed_corpus)
    doc_ll, phi_ss = lda.e_step()
    # phi_ss is zero outside the columns of terms that occur: store those columns only
    cols = numpy.unique(ids)
    mask = numpy.ones(phi_ss.shape[1], dtype=bool)
    mask[cols] = False
    assert not phi_ss[:, mask].any()
    out = dict(K=K, V=lda._number_of_types, row_ptr=row_ptr, ids=ids, cts=cts,
               alpha=alpha0, gamma=lda._gamma.copy(), phi_cols=cols, phi_ss_cols=phi_ss[:, cols],
               doc_ll=numpy.float64(doc_ll),
               eta_sha1=numpy.array(hashlib.sha1(eta0.tobytes()).hexdigest()))
    if em_warm == 0:
        # eta0 is the Gamma(100, 1/100) draw of variational_bayes.py:95 from seed 0: regenerate, don't store
        from pylda_b200 import synthetic
        assert numpy.array_equal(synthetic.initial_eta(K, out["V"], 0), eta0)
        out["eta_seed"] = 0
    else:
        out["eta"] = eta0
    if heldout_docs is not None:
        parsed = lda.parse_data(heldout_docs)
        h_row_ptr, h_ids, h_cts = O.csr_from_parsed(*parsed)
        words_ll, h_gamma = lda.e_step(parsed)
        out.update(h_row_ptr=h_row_ptr, h_ids=h_ids, h_cts=h_cts,
                   h_words_ll=numpy.float64(words_ll), h_gamma=h_gamma)
    # pin the restatement right here too
    r = O.e_step(row_ptr, ids, cts, eta0, alpha0)
    rel = lambda a, b: float(numpy.max(numpy.abs(a - b) / numpy.maximum(numpy.abs(b), 1e-300)))
    print("%-16s D=%d V=%d K=%d nnz=%d  oracle-vs-reference: gamma %.2e phi_ss(abs) %.2e doc_ll %.2e" % (
        name, len(row_ptr) - 1, out["V"], K, len(ids), rel(r["gamma"], out["gamma"]),
        float(numpy.max(numpy.abs(r["phi_ss"] - phi_ss))), abs(r["doc_ll"] - doc_ll) / abs(doc_ll)))
    numpy.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    return lda


def main():
    os.makedirs(GOLD, exist_ok=True)
    inf_mod, vb_mod = ref_shim.load()

    ap = _extract("associated-press.tar.gz", "associated-press")
    ap_train = _read_docs(os.path.join(ap, "train.dat"))
    ap_test = _read_docs(os.path.join(ap, "test.dat"))
    ap_vocab = _read_vocab(os.path.join(ap, "voc.dat"))

    small = "--only-nips-full" not in sys.argv
    if small:
        # config 1 shape, first 200 docs, EM iteration 1 (random eta) + held-out branch
        _run_case("ap200_k10", vb_mod, ap_train[:200], ap_vocab, 10, heldout_docs=ap_test[:40])
        # same after 3 EM iterations (warm eta, asymmetric alpha, early-exit documents)
        _run_case("ap200_k10_warm3", vb_mod, ap_train[:200], ap_vocab, 10, em_warm=3)

    # config 4 shape: nips.88-05 has doc.dat only
    nips = _extract("parsed/nips.88-05.tar.gz", "nips.88-05")
    nips_docs = _read_docs(os.path.join(nips, "doc.dat"))
    nips_vocab = _read_vocab(os.path.join(nips, "voc.dat"))
    from pylda_b200 import synthetic
    if small:
        _run_case("nips24_k200", vb_mod, nips_docs[:24], nips_vocab, 200)

        # synthetic, config 2 shape scaled down (K=50), rendered as text so it goes through
        # the reference's own parse_data
        row_ptr, ids, cts = synthetic.synthetic_corpus(96, 1500, seed=1236, length="poisson", mean_len=100)
        docs = synthetic.render_text(row_ptr, ids, cts)
        vocab = ["w%d" % i for i in range(1500)]
        _run_case("syn96_k50", vb_mod, docs, vocab, 50, em_warm=1)
        # zipf lengths incl. one long document (config 3 shape scaled down, K=100)
        row_ptr, ids, cts = synthetic.synthetic_corpus(48, 3000, seed=1237, length="zipf")
        docs = synthetic.render_text(row_ptr, ids, cts)
        vocab = ["w%d" % i for i in range(3000)]
        _run_case("zipf48_k100", vb_mod, docs, vocab, 100)

    if "--nips-full" in sys.argv:
        # config 4 in full: nips.88-05 (doc.dat staged as the training corpus), K = 200, EM iteration 1.
        # Kept small: the CSR as int16/uint16, gamma in full, phi_ss as every 16th occurring column plus its
        # row and column sums, the ELBO, and the trip counts of the restatement (pinned to the reference here).
        K = 200
        numpy.random.seed(0)
        lda = vb_mod.VariationalBayes()
        lda._initialize(nips_docs, nips_vocab, K, 1.0 / K, 1.0 / len(set(nips_vocab)))
        eta0 = lda._eta.copy()
        from pylda_b200 import synthetic
        assert numpy.array_equal(synthetic.initial_eta(K, lda._number_of_types, 0), eta0)
        row_ptr, ids, cts = O.csr_from_parsed(*lda._parsed_corpus)
        doc_ll, phi_ss = lda.e_step()
        r = O.e_step(row_ptr, ids, cts, eta0, lda._alpha_alpha, return_iters=True)
        rel = lambda a, b: float(numpy.max(numpy.abs(a - b) / numpy.maximum(numpy.abs(b), 1e-300)))
        print("nips_full_k200 D=%d V=%d nnz=%d  oracle-vs-reference: gamma %.2e phi_ss(abs) %.2e doc_ll %.2e  mean trips %.1f" % (
            len(row_ptr) - 1, lda._number_of_types, len(ids), rel(r["gamma"], lda._gamma),
            float(numpy.max(numpy.abs(r["phi_ss"] - phi_ss))), abs(r["doc_ll"] - doc_ll) / abs(doc_ll), r["iters"].mean()))
        assert ids.max() < 32768 and cts.max() < 65536
        cols = numpy.unique(ids)[::16]
        numpy.savez_compressed(os.path.join(GOLD, "nips_full_k200.npz"), K=K, V=lda._number_of_types, eta_seed=0,
                               row_ptr=row_ptr.astype(numpy.int32), ids=ids.astype(numpy.int16), cts=cts.astype(numpy.uint16),
                               alpha=lda._alpha_alpha.copy(), gamma=lda._gamma.copy(), phi_cols=cols,
                               phi_ss_cols=phi_ss[:, cols], phi_rowsum=phi_ss.sum(axis=1), phi_colsum=phi_ss.sum(axis=0),
                               doc_ll=numpy.float64(doc_ll), iters=r["iters"].astype(numpy.int8),
                               eta_sha1=numpy.array(hashlib.sha1(eta0.tobytes()).hexdigest()))

    if "--trace" in sys.argv:
        # config 1 in full: AP, K=10, 20 VB iterations through the reference's learning()
        numpy.random.seed(0)
        lda = vb_mod.VariationalBayes()
        lda._initialize(ap_train, ap_vocab, 10, 1.0 / 10, 1.0 / len(ap_vocab))
        row_ptr, ids, cts = O.csr_from_parsed(*lda._parsed_corpus)
        eta0 = lda._eta.copy()
        alpha0 = lda._alpha_alpha.copy()
        elbo, sum_gamma, sum_alpha = [], [], []
        n_it = 20
        for it in range(n_it):
            elbo.append(lda.learning())
            sum_gamma.append(lda._gamma.sum())
            sum_alpha.append(lda._alpha_alpha.sum())
        numpy.savez_compressed(os.path.join(GOLD, "ap_full_k10_trace.npz"), K=10, V=lda._number_of_types,
                               row_ptr=row_ptr, ids=ids, cts=cts, alpha=alpha0,
                               eta_sha1=numpy.array(hashlib.sha1(eta0.tobytes()).hexdigest()),
                               eta_seed=0, alpha_beta=lda._alpha_beta[0],
                               elbo=numpy.array(elbo), sum_gamma=numpy.array(sum_gamma),
                               sum_alpha=numpy.array(sum_alpha), final_alpha=lda._alpha_alpha.copy(),
                               final_gamma_rowsum=lda._gamma.sum(axis=1))
        print("trace", elbo[:3], "...", elbo[-1])


if __name__ == "__main__":
    main()
