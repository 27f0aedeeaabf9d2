"""The drop-in class surface on the GPU: VariationalBayes.learning / inference / pickling and the
launch_train / launch_test drivers, against fixtures generated from the reference itself
(tests/golden/*.npz, oracle/make_golden.py).  Tolerance 1e-5 relative (north_star)."""
import hashlib
import os
import pickle

import numpy
import pytest

from tests.util import GOLDEN, RTOL, load_golden, max_rel

pytestmark = pytest.mark.gpu


def _vb_from_csr(row_ptr, ids, cts, K, V, eta, alpha, alpha_beta):
    """A VariationalBayes in the state _initialize (variational_bayes.py:82-95) leaves it in, but with a
    given parsed corpus / eta instead of text and a fresh RNG draw (type ids come from set() order,
    so text cannot reproduce a fixture's ids)."""
    from oracle import estep_oracle as O          # checker-side helper: CSR -> the reference's list type
    from pylda_b200 import variational_bayes as vb
    lda = vb.VariationalBayes()
    lda._type_to_index = {"w%d" % i: i for i in range(V)}
    lda._index_to_type = {i: "w%d" % i for i in range(V)}
    lda._vocab = list(lda._type_to_index.keys())
    lda._number_of_types = V
    lda._counter = 0
    lda._number_of_topics = K
    lda._alpha_alpha = numpy.array(alpha, dtype=numpy.float64)
    lda._alpha_beta = numpy.zeros(V) + alpha_beta
    lda._parsed_corpus = O.parsed_from_csr(row_ptr, ids, cts)
    lda._number_of_documents = len(row_ptr) - 1
    lda._gamma = numpy.zeros((lda._number_of_documents, K)) + lda._alpha_alpha[numpy.newaxis, :] + 1.0 * V / K
    lda._eta = numpy.array(eta)
    from pylda_b200.variational_bayes import pack_parsed_corpus
    lda._train_csr = pack_parsed_corpus(lda._parsed_corpus)
    lda._train_uploaded = False
    return lda


@pytest.mark.parametrize("resident", ["1", "0"])
def test_learning_matches_reference_trace_config1(resident, monkeypatch):
    """BASELINE.json configs[0]: associated-press, K=10, 20 VB iterations against the reference's own
    ELBO trace.  resident=1: the model stays in HBM (device E-step + device
    M-step + device alpha statistics, alpha Newton update on the host); resident=0: E-step on the GPU,
    M-step on the host exactly as the reference."""
    from pylda_b200 import synthetic
    monkeypatch.setenv("PYLDA_RESIDENT", resident)
    z = numpy.load(os.path.join(GOLDEN, "ap_full_k10_trace.npz"))
    K, V = int(z["K"]), int(z["V"])
    eta0 = synthetic.initial_eta(K, V, int(z["eta_seed"]))
    assert hashlib.sha1(eta0.tobytes()).hexdigest() == str(z["eta_sha1"])
    lda = _vb_from_csr(z["row_ptr"], z["ids"], z["cts"], K, V, eta0, z["alpha"], float(z["alpha_beta"]))
    elbo, sum_gamma, sum_alpha = [], [], []
    for _ in range(len(z["elbo"])):
        elbo.append(lda.learning())
        sum_gamma.append(lda._gamma.sum())
        sum_alpha.append(lda._alpha_alpha.sum())
    rel = numpy.abs(numpy.array(elbo) - z["elbo"]) / numpy.abs(z["elbo"])
    print("ELBO trace max rel err %.2e; final ELBO %.10f (reference %.10f)" % (rel.max(), elbo[-1], z["elbo"][-1]))
    assert rel.max() <= RTOL
    assert numpy.allclose(sum_gamma, z["sum_gamma"], rtol=RTOL, atol=0)
    assert numpy.allclose(sum_alpha, z["sum_alpha"], rtol=RTOL, atol=0)
    assert max_rel(lda._alpha_alpha, z["final_alpha"]) <= RTOL
    assert max_rel(lda._gamma.sum(axis=1), z["final_gamma_rowsum"]) <= RTOL
    assert abs(lda._eta.sum() - (float(z["cts"].sum()) + K * V * float(z["alpha_beta"]))) <= 1e-6 * z["cts"].sum()
    assert bool(lda.__dict__.get("_model_on_device")) == (resident == "1")


def test_inference_heldout_and_gamma_untouched():
    g = load_golden("ap200_k10")
    lda = _vb_from_csr(g["row_ptr"], g["ids"], g["cts"], g["K"], g["V"], g["eta"], g["alpha"], 1.0 / g["V"])
    from oracle import estep_oracle as O
    before = lda._gamma.copy()
    words_ll, gam = lda.e_step(O.parsed_from_csr(g["h_row_ptr"], g["h_ids"], g["h_cts"]))
    assert numpy.array_equal(lda._gamma, before)                       # :216 -- held-out leaves self._gamma alone
    assert abs(words_ll - float(g["h_words_ll"])) <= RTOL * abs(float(g["h_words_ll"]))
    assert max_rel(gam, g["h_gamma"]) <= RTOL
    doc_ll, phi = lda.e_step()
    assert abs(doc_ll - g["doc_ll"]) <= RTOL * abs(g["doc_ll"])
    assert max_rel(lda._gamma, g["gamma"]) <= RTOL
    assert phi.shape == (g["K"], g["V"])


def test_pickle_roundtrip_keeps_working():
    """launch_train.py:203-204 pickles the whole inferencer; launch_test.py:92 loads it and runs inference."""
    g = load_golden("syn96_k50")
    lda = _vb_from_csr(g["row_ptr"], g["ids"], g["cts"], g["K"], g["V"], g["eta"], g["alpha"], 1.0 / g["V"])
    lda.learning()
    blob = pickle.dumps(lda)
    assert b"c_void_p" not in blob
    twin = pickle.loads(blob)
    assert twin._native is None and twin._counter == 1
    a = lda.learning()
    b = twin.learning()                                                 # a second device context, same state
    assert abs(a - b) <= 1e-9 * abs(a)


def test_launch_train_and_test_drivers(tmp_path, capsys):
    """The reference's CLI surface end to end on a small rendered corpus: option.txt, snapshots, model
    pickle, held-out evaluation file (launch_train.py:126-204, launch_test.py:60-97)."""
    from pylda_b200 import launch_test, launch_train, synthetic
    V, K = 300, 5
    row_ptr, ids, cts = synthetic.synthetic_corpus(60, V, seed=8, length="poisson", mean_len=40)
    docs = synthetic.render_text(row_ptr, ids, cts)
    corpus = tmp_path / "toy"
    corpus.mkdir()
    (corpus / "train.dat").write_text("\n".join(docs[:50]) + "\n")
    (corpus / "test.dat").write_text("\n".join(docs[50:]) + "\n")
    (corpus / "voc.dat").write_text("".join("w%d\t1\n" % i for i in range(V)))
    out = tmp_path / "out"
    numpy.random.seed(3)
    run_dir = launch_train.main(["--input_directory=%s/" % corpus, "--output_directory=%s" % out,
                                 "--number_of_topics=%d" % K, "--training_iterations=4", "--snapshot_interval=2",
                                 "--inference_mode=2"])
    assert run_dir is not None and os.path.isdir(run_dir)
    names = sorted(os.listdir(run_dir))
    assert names == ["exp_beta-2", "exp_beta-4", "exp_gamma-2", "exp_gamma-4", "model-4", "option.txt"]
    opts = dict(line.strip().split("=", 1) for line in open(os.path.join(run_dir, "option.txt")))
    assert opts["number_of_topics"] == str(K) and opts["inference_mode"] == "2" and opts["corpus_name"] == "toy"
    assert float(opts["alpha_alpha"]) == 1.0 / K and abs(float(opts["alpha_beta"]) - 1.0 / V) < 1e-15
    text = capsys.readouterr().out
    assert text.count("e_step and m_step of iteration") == 4
    beta_lines = open(os.path.join(run_dir, "exp_beta-4")).read().split("\n")
    assert beta_lines[0] == "==========\t0\t==========" and len(beta_lines) == K * (V + 1) + 1
    assert len(open(os.path.join(run_dir, "exp_gamma-4")).read().strip().split("\n")) == 50
    # unknown / out-of-scope modes are reported, not run (launch_train.py:190-192)
    import time
    time.sleep(1.1)          # the run directory name has one-second resolution (launch_train.py:127)
    assert launch_train.main(["--input_directory=%s" % corpus, "--output_directory=%s" % out,
                              "--number_of_topics=%d" % K, "--training_iterations=1", "--inference_mode=0"]) is None
    res = launch_test.main(["--input_directory=%s" % corpus, "--model_directory=%s" % run_dir, "--snapshot_index=4"])
    assert 4 in res
    ll, gam = res[4]
    assert numpy.isfinite(ll) and ll < 0 and gam.shape == (10, K)
    saved = numpy.loadtxt(os.path.join(run_dir, "test-4"))
    assert numpy.allclose(saved, gam, rtol=1e-12)


def test_device_top_words_match_host_export(ctx, tmp_path):
    """pylda_top_words (device side of export_beta, variational_bayes.py:326-341) against the reference's host
    formula: per topic exp(E_log_eta - logsumexp) sorted descending; and the exported file is the same text."""
    import os
    import scipy.special
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    K, V = 30, 900
    rs = numpy.random.RandomState(5)
    eta = rs.gamma(0.3, 1.0, (K, V)) + 1e-3
    ctx.set_model(eta, numpy.full(K, 0.1))
    idx, prob = ctx.top_words(25)
    E = O.compute_dirichlet_expectation(eta)
    beta = numpy.exp(E - scipy.special.logsumexp(E, axis=1)[:, None])
    order = numpy.argsort(-beta, axis=1)[:, :25]
    assert numpy.allclose(prob, numpy.take_along_axis(beta, order, axis=1), rtol=1e-12)
    assert numpy.array_equal(idx, order)          # (continuous random eta: no ties)
    full_idx, full_prob = ctx.top_words(V)
    assert numpy.array_equal(numpy.sort(full_idx, axis=1), numpy.tile(numpy.arange(V), (K, 1)))
    assert numpy.all(numpy.diff(full_prob, axis=1) <= 0) and numpy.allclose(full_prob.sum(axis=1), 1.0, rtol=1e-12)
