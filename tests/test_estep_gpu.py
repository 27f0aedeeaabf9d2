"""GPU parity tests proper: CUDA path through the C ABI vs (a) fixtures generated from the
reference itself and (b) the oracle on seeded inputs.  Tolerance: 1e-5 relative (north_star)."""
import numpy
import pytest

from tests.util import CASES, PHI_FLOOR, RTOL, load_golden, max_rel

pytestmark = pytest.mark.gpu


def _check(out, ref_gamma, ref_phi, ref_doc_ll, tag):
    eg = max_rel(out["gamma"], ref_gamma)
    ep = max_rel(out["phi_ss"], ref_phi, floor=PHI_FLOOR)
    el = abs(out["doc_ll"] - ref_doc_ll) / abs(ref_doc_ll)
    print("%s: gamma %.2e phi_ss %.2e doc_ll %.2e" % (tag, eg, ep, el))
    if out.get("stats"):
        assert out["stats"]["revived_docs"] == 0, (tag, "a topic eliminated as dead came back")
    assert eg <= RTOL, (tag, "gamma", eg)
    assert ep <= RTOL, (tag, "phi_ss", ep)
    assert el <= RTOL, (tag, "doc_ll", el)
    return eg, ep, el


@pytest.mark.parametrize("name", CASES)
def test_golden_train_branch(ctx, name):
    g = load_golden(name)
    ctx.set_corpus(0, g["row_ptr"], g["ids"], g["cts"])
    out = ctx.estep(0, g["eta"], g["alpha"], 50, 1e-6)
    _check(out, g["gamma"], g["phi_ss"], g["doc_ll"], name)
    # invariants: each phi row sums to 1
    N = numpy.add.reduceat(g["cts"].astype(numpy.float64), g["row_ptr"][:-1])
    assert numpy.allclose(out["gamma"].sum(axis=1), g["alpha"].sum() + N, rtol=1e-12)
    assert abs(out["phi_ss"].sum() - g["cts"].sum()) <= 1e-9 * g["cts"].sum()


def test_golden_heldout_branch(ctx):
    g = load_golden("ap200_k10")
    ctx.set_corpus(1, g["h_row_ptr"], g["h_ids"], g["h_cts"])
    out = ctx.estep(1, g["eta"], g["alpha"], 50, 1e-6, heldout=True, want_phi=False)
    eg = max_rel(out["gamma"], g["h_gamma"])
    el = abs(out["words_ll"] - float(g["h_words_ll"])) / abs(float(g["h_words_ll"]))
    print("heldout: gamma %.2e words_ll %.2e" % (eg, el))
    assert eg <= RTOL and el <= RTOL


def test_corpus_roundtrip_bit_exact(ctx):
    g = load_golden("zipf48_k100")
    ctx.set_corpus(0, g["row_ptr"], g["ids"], g["cts"])
    r, i, c = ctx.get_corpus(0)
    assert numpy.array_equal(r, g["row_ptr"]) and numpy.array_equal(i, g["ids"]) and numpy.array_equal(c, g["cts"])


@pytest.mark.parametrize("K,V,D,length", [
    (3, 50, 40, "poisson"), (10, 300, 64, "poisson"), (16, 300, 64, "zipf"), (25, 400, 48, "poisson"),
    (50, 2000, 64, "zipf"), (64, 500, 32, "poisson"), (100, 5000, 96, "zipf"), (128, 800, 24, "poisson"),
    (200, 1500, 24, "zipf"), (500, 3000, 12, "poisson"), (501, 1200, 6, "poisson"), (1000, 1500, 4, "poisson"),
])
def test_against_oracle_shapes(ctx, K, V, D, length):
    """Every lane shape (LK,J), every group size class and the streaming path (zipf lengths reach
    documents whose tile does not fit in shared memory for the larger K)."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    row_ptr, ids, cts = synthetic.synthetic_corpus(D, V, seed=77 + K, length=length, mean_len=60)
    eta = synthetic.initial_eta(K, V, seed=K)
    rs = numpy.random.RandomState(K)
    alpha = rs.uniform(0.02, 0.5, K) if K % 2 else numpy.full(K, 1.0 / K)
    ref = O.e_step(row_ptr, ids, cts, eta, alpha, 50, 1e-6, heldout=True, return_iters=True)
    ctx.set_corpus(0, row_ptr, ids, cts)
    out = ctx.estep(0, eta, alpha, 50, 1e-6, heldout=True)
    _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "K=%d" % K)
    assert abs(out["words_ll"] - ref["words_ll"]) <= RTOL * abs(ref["words_ll"])
    res = ctx.get_results(0, gamma=False, phi=False, iters=True)
    assert numpy.array_equal(res["iters"], ref["iters"])
    print("stats", out["stats"])


@pytest.mark.parametrize("kernel", ["v2", "hybrid", "default"])
def test_every_kernel_generation_matches_oracle(ctx, kernel, monkeypatch):
    """The library holds several per-document kernels (shared-memory tile, register tile + narrow stages + streaming
    [default], hybrid register / shared-memory thread-block cluster for long documents); PYLDA_KERNEL selects one.
    All of them must agree with the oracle on a corpus with short, medium and very long documents."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    K, V = 100, 4000
    a = synthetic.synthetic_corpus(120, V, seed=99, length="zipf")
    b = synthetic.synthetic_corpus(6, V, seed=98, length="poisson", mean_len=5000)   # ~1000 unique terms each
    row_ptr = numpy.concatenate([a[0], a[0][-1] + b[0][1:]])
    ids, cts = numpy.concatenate([a[1], b[1]]), numpy.concatenate([a[2], b[2]])
    assert numpy.diff(row_ptr).max() > 600          # long documents: cluster / streaming paths
    eta = synthetic.initial_eta(K, V, seed=1)
    alpha = numpy.full(K, 1.0 / K)
    ref = O.e_step(row_ptr, ids, cts, eta, alpha, 50, 1e-6, return_iters=True)
    if kernel != "default":
        monkeypatch.setenv("PYLDA_KERNEL", kernel)
    ctx.set_corpus(0, row_ptr, ids, cts)
    out = ctx.estep(0, eta, alpha, 50, 1e-6)
    _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "kernel=%s" % kernel)
    it = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
    assert numpy.array_equal(it, ref["iters"])
    print("stats", out["stats"])


def test_empty_and_single_term_documents(ctx):
    """Ragged edge cases: documents with no terms (n_d = 0), one term, and a corpus of one document."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    K, V = 12, 50
    row_ptr = numpy.array([0, 0, 1, 1, 4, 4], dtype=numpy.int64)      # docs: empty, 1 term, empty, 3 terms, empty
    ids = numpy.array([7, 3, 9, 49], dtype=numpy.int32)
    cts = numpy.array([5, 1, 2, 1], dtype=numpy.int32)
    eta = synthetic.initial_eta(K, V, seed=2)
    alpha = numpy.linspace(0.05, 0.4, K)
    ref = O.e_step(row_ptr, ids, cts, eta, alpha, 50, 1e-6)
    ctx.set_corpus(0, row_ptr, ids, cts)
    out = ctx.estep(0, eta, alpha, 50, 1e-6)
    _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "ragged")
    ctx.set_corpus(0, row_ptr[3:5] - 1, ids[1:], cts[1:])              # a single document
    ref1 = O.e_step(row_ptr[3:5] - 1, ids[1:], cts[1:], eta, alpha, 50, 1e-6)
    out1 = ctx.estep(0, eta, alpha, 50, 1e-6)
    _check(out1, ref1["gamma"], ref1["phi_ss"], ref1["doc_ll"], "single")


def test_warm_model_early_exit(ctx):
    """After a few EM iterations documents converge before the cap: trip counts must match the oracle's."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    K, V = 20, 600
    row_ptr, ids, cts = synthetic.synthetic_corpus(80, V, seed=5, length="poisson", mean_len=80)
    eta = synthetic.initial_eta(K, V, 0)
    alpha = numpy.full(K, 1.0 / K)
    ab = numpy.full(V, 1.0 / V)
    trace, eta_w, alpha_w, _ = O.learning_trace(row_ptr, ids, cts, eta, alpha, ab, 4)
    ref = O.e_step(row_ptr, ids, cts, eta_w, alpha_w, return_iters=True)
    ctx.set_corpus(0, row_ptr, ids, cts)
    out = ctx.estep(0, eta_w, alpha_w)
    _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "warm")
    it = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
    assert (ref["iters"] < 50).any()
    assert numpy.mean(it == ref["iters"]) >= 0.98   # a borderline |d gamma| == tol document may flip by one trip


def test_max_iter_and_tol_arguments(ctx):
    from oracle import estep_oracle as O
    g = load_golden("syn96_k50")
    ctx.set_corpus(0, g["row_ptr"], g["ids"], g["cts"])
    for max_iter, tol in ((1, 1e-6), (7, 1e-6), (50, 1e-2)):
        ref = O.e_step(g["row_ptr"], g["ids"], g["cts"], g["eta"], g["alpha"], max_iter, tol)
        out = ctx.estep(0, g["eta"], g["alpha"], max_iter, tol)
        _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "max_iter=%d tol=%g" % (max_iter, tol))


def test_special_functions_vs_scipy(ctx):
    import scipy.special as sp
    rs = numpy.random.RandomState(1)
    x = numpy.concatenate([10 ** rs.uniform(-6, 7, 200000), rs.uniform(0, 12, 100000), [1e-6, 1.0, 2.0, 1e7]])
    psi = ctx.special("digamma", x)
    ref = sp.psi(x)
    err = numpy.abs(psi - ref) / numpy.maximum(1.0, numpy.abs(ref))
    print("digamma max err", err.max())
    assert err.max() <= 1e-13
    ex = ctx.special("exp_digamma", x)
    refe = numpy.exp(ref)
    ok = refe > 1e-290
    erre = numpy.abs(ex[ok] - refe[ok]) / refe[ok] / numpy.maximum(1.0, numpy.abs(ref[ok]))
    print("exp(digamma) max err / cond", erre.max())
    assert erre.max() <= 1e-13
    ex2 = ctx.special("exp_digamma_v2", x)
    erre2 = numpy.abs(ex2[ok] - refe[ok]) / refe[ok] / numpy.maximum(1.0, numpy.abs(ref[ok]))
    print("exp(digamma) v2 max err / cond", erre2.max())
    assert erre2.max() <= 1e-13
    rc = ctx.special("rcp", x)
    errr = numpy.abs(rc * x - 1.0)
    print("rcp max err", errr.max())
    assert errr.max() <= 1e-15
    lg = ctx.special("lgamma", x)
    refl = sp.gammaln(x)
    errl = numpy.abs(lg - refl) / numpy.maximum(1.0, numpy.abs(refl))
    assert errl.max() <= 1e-13


def test_dirichlet_expectation_vs_oracle(ctx):
    from oracle import estep_oracle as O
    rs = numpy.random.RandomState(3)
    eta = numpy.concatenate([rs.gamma(100., 0.01, (7, 900)), rs.gamma(0.01, 1.0, (7, 900)) + 1e-4], axis=0)
    out = ctx.dirichlet_expectation(eta)
    ref = O.compute_dirichlet_expectation(eta)
    assert max_rel(out, ref, floor=1.0) <= 1e-12


def test_error_behaviour(ctx):
    g = load_golden("syn96_k50")
    ctx.set_corpus(0, g["row_ptr"], g["ids"], g["cts"])
    with pytest.raises(RuntimeError):
        ctx.estep(0, g["eta"][:, :10], g["alpha"])            # V smaller than the largest term id
    with pytest.raises(RuntimeError):
        ctx.estep(0, g["eta"], g["alpha"], max_iter=0)
    with pytest.raises(RuntimeError):
        bad = g["alpha"].copy(); bad[0] = -1.0
        ctx.estep(0, g["eta"], bad)
    with pytest.raises(RuntimeError):
        ctx.set_corpus(0, g["row_ptr"], g["ids"], numpy.zeros_like(g["cts"]))   # counts must be >= 1


def test_resident_em_loop_matches_oracle(ctx):
    """eta stays on the device across EM iterations -- E-step, device M-step
    (variational_bayes.py:222-226) and the alpha statistics (:232-233) computed from the resident gamma.
    Three EM iterations against the oracle's learning loop with alpha held fixed."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    K, V = 30, 900
    row_ptr, ids, cts = synthetic.synthetic_corpus(120, V, seed=12, length="poisson", mean_len=70)
    eta = synthetic.initial_eta(K, V, 0)
    alpha = numpy.full(K, 1.0 / K)
    alpha_beta = 1.0 / V
    ctx.set_corpus(0, row_ptr, ids, cts)
    ctx.set_model(eta, alpha)
    for em in range(3):
        ref = O.e_step(row_ptr, ids, cts, eta, alpha, 50, 1e-6)
        topic_ref, eta_new, alpha_ss_ref = O.m_step(eta, ref["gamma"], ref["phi_ss"], numpy.full(V, alpha_beta))
        ctx.estep_resident(0, 50, 1e-6, want_alpha_ss=True)
        res = ctx.get_results(0, gamma=True, phi=False, alpha_ss=True)
        topic_ll, eta_dev = ctx.mstep_resident(alpha_beta, want_eta=True)
        assert max_rel(res["gamma"], ref["gamma"]) <= RTOL
        assert abs(res["doc_ll"] - ref["doc_ll"]) <= RTOL * abs(ref["doc_ll"])
        assert max_rel(res["alpha_ss"], alpha_ss_ref) <= RTOL
        assert abs(topic_ll - topic_ref) <= RTOL * abs(topic_ref)
        assert max_rel(eta_dev, eta_new) <= RTOL
        eta = eta_new


def test_page_locked_gamma_buffer(ctx, monkeypatch):
    """pylda_estep with a page-locked gamma buffer.  Without the hand-over to the narrow stages the kernels store
    gamma straight into host memory (no D x K copy at the end); with it (default) gamma leaves by DMA, and the copy
    starts before the kernels of the long documents (which run last), whose rows follow through the buffer's device
    alias.  Either way the results are identical to the pageable path."""
    g = load_golden("zipf48_k100")
    ctx.set_corpus(0, g["row_ptr"], g["ids"], g["cts"])
    plain = ctx.estep(0, g["eta"], g["alpha"], 50, 1e-6)
    D = len(g["row_ptr"]) - 1
    n_late = int((numpy.diff(g["row_ptr"]) > 272).sum())     # longer than every register / shared-memory class at K = 100
    assert 0 < n_late < D
    pinned = numpy.full((D, g["K"]), -1.0)
    ctx.pin(pinned)
    try:
        out = ctx.estep(0, g["eta"], g["alpha"], 50, 1e-6, gamma_out=pinned)
        assert out["gamma"] is pinned and out["stats"]["docs_narrow"] > 0
        assert 0 < out["stats"]["gamma_rows_early"] <= D - n_late
        assert numpy.array_equal(pinned, plain["gamma"])
        assert numpy.array_equal(ctx.get_results(0, gamma=True, phi=False)["gamma"], plain["gamma"])   # still on the device
        pinned.fill(-1.0)
        both = ctx.estep(0, g["eta"], g["alpha"], 50, 1e-6, gamma_out=pinned, want_alpha_ss=True)
        assert both["stats"]["gamma_rows_early"] > 0 and numpy.array_equal(pinned, plain["gamma"])
        monkeypatch.setenv("PYLDA_EARLY_COPY", "0")
        pinned.fill(-1.0)
        late = ctx.estep(0, g["eta"], g["alpha"], 50, 1e-6, gamma_out=pinned)
        assert late["stats"]["gamma_rows_early"] == 0 and numpy.array_equal(pinned, plain["gamma"])
        monkeypatch.delenv("PYLDA_EARLY_COPY")
        monkeypatch.setenv("PYLDA_PARK", "0")
        pinned.fill(-1.0)
        out = ctx.estep(0, g["eta"], g["alpha"], 50, 1e-6, gamma_out=pinned)
        assert out["gamma"] is pinned and out["stats"]["docs_narrow"] == 0
        assert max_rel(pinned, plain["gamma"]) <= 1e-11
        assert out["doc_ll"] == plain["doc_ll"] or abs(out["doc_ll"] - plain["doc_ll"]) <= 1e-12 * abs(plain["doc_ll"])
        with pytest.raises(RuntimeError):
            ctx.get_results(0, gamma=True, phi=False)        # gamma of that call was written in place, not on the device
        again = ctx.estep(0, g["eta"], g["alpha"], 50, 1e-6, gamma_out=pinned, want_alpha_ss=True)   # falls back to the copy
        assert max_rel(again["gamma"], plain["gamma"]) <= 1e-11
    finally:
        ctx.unpin(pinned)


def test_long_documents_finish_in_the_compact_stage(ctx, monkeypatch):
    """Documents of more than 192 terms are handed over by estep_stream / estep_v2 once at most 32 topics are alive
    and finished by estep_longc on a compact tile (shared memory or, for the longest, global scratch).  A corpus
    drawn from an LDA model with (nearly) its own topics as the model, so that the long documents get there: same
    gamma, statistics, ELBO and trip counts as the oracle, and as the library with that stage switched off."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    K, V, D = 100, 3000, 160
    row_ptr, ids, cts = synthetic.lda_corpus(D, V, seed=7)
    n = numpy.diff(row_ptr)
    assert (n > 400).sum() >= 3 and ((n > 192) & (n <= 272)).sum() >= 1
    rng = numpy.random.default_rng(7)
    topics = rng.gamma(0.05, 1.0, size=(50, V))                        # the generator's first draw: its topics
    topics /= topics.sum(axis=1, keepdims=True)
    eta = numpy.concatenate([0.01 + 4000.0 * topics, 0.01 + 0.02 * numpy.random.RandomState(3).rand(K - 50, V)])
    alpha = numpy.full(K, 1.0 / K)
    ref = O.e_step(row_ptr, ids, cts, eta, alpha, 50, 1e-6, return_iters=True)
    ctx.set_corpus(0, row_ptr, ids, cts)
    out = ctx.estep(0, eta, alpha, 50, 1e-6)
    it = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
    print("stats", out["stats"])
    assert out["stats"]["docs_long_compact"] >= 5 and out["stats"]["revived_docs"] == 0
    _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "long compact")
    assert numpy.array_equal(it, ref["iters"])
    monkeypatch.setenv("PYLDA_LONGC_SMEM", "0")                        # every tile through the global scratch
    scr = ctx.estep(0, eta, alpha, 50, 1e-6)
    assert scr["stats"]["docs_long_compact"] == out["stats"]["docs_long_compact"]
    assert numpy.array_equal(scr["gamma"], out["gamma"])
    assert max_rel(scr["phi_ss"], out["phi_ss"], floor=PHI_FLOOR) <= 1e-11
    monkeypatch.delenv("PYLDA_LONGC_SMEM")
    # the streaming kernel's hand-over is only used for a peaked model (flatness statistic of k_build_B): forced off,
    # only the shared-memory class hands over -- fewer documents in the compact stage, same results
    monkeypatch.setenv("PYLDA_FLAT_THRESHOLD", "0")
    lean = ctx.estep(0, eta, alpha, 50, 1e-6)
    assert 0 < lean["stats"]["docs_long_compact"] < out["stats"]["docs_long_compact"]
    assert max_rel(lean["gamma"], out["gamma"]) <= 1e-11
    assert abs(lean["doc_ll"] - out["doc_ll"]) <= 1e-12 * abs(out["doc_ll"])
    monkeypatch.delenv("PYLDA_FLAT_THRESHOLD")
    monkeypatch.setenv("PYLDA_PARK_LONG", "0")
    off = ctx.estep(0, eta, alpha, 50, 1e-6)
    assert off["stats"]["docs_long_compact"] == 0
    assert max_rel(off["gamma"], out["gamma"]) <= 1e-11
    assert max_rel(off["phi_ss"], out["phi_ss"], floor=PHI_FLOOR) <= 1e-10
    assert abs(off["doc_ll"] - out["doc_ll"]) <= 1e-12 * abs(out["doc_ll"])
    assert numpy.array_equal(ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"], it)


@pytest.mark.parametrize("K", [4, 8, 20, 33])
def test_few_topics_tiny_alpha_every_document_is_handed_over(ctx, K):
    """K at or below the hand-over thresholds with an alpha small enough for the elimination to be on (alpha =
    0.004): every document has <= 32 live topics from the first trip on.  Whatever each kernel then does with it
    (the streaming kernel and estep_v2 hand long documents to estep_longc at once, on a compact tile wider than K;
    the register-tile kernel only compacts when K > 32), the results are the oracle's.  Cold and peaked models."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    V = 3000
    row_ptr, ids, cts = synthetic.lda_corpus(300, V, seed=11, topics=6)
    n = numpy.diff(row_ptr)
    assert (n > 192).sum() >= 3 and (n <= 24).sum() >= 3
    alpha = numpy.full(K, 0.004)
    rng = numpy.random.default_rng(11)
    topics = rng.gamma(0.05, 1.0, size=(6, V))
    topics /= topics.sum(axis=1, keepdims=True)
    peaked = 0.01 + 3000.0 * topics[numpy.arange(K) % 6] * (1.0 + 0.3 * numpy.random.RandomState(5).rand(K, V))
    ctx.set_corpus(0, row_ptr, ids, cts)
    for tag, eta in (("cold", synthetic.initial_eta(K, V, 3)), ("peaked", peaked)):
        ref = O.e_step(row_ptr, ids, cts, eta, alpha, 50, 1e-6, return_iters=True)
        out = ctx.estep(0, eta, alpha, 50, 1e-6)
        it = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
        print(tag, "K=%d" % K, "stats", out["stats"])
        _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "K=%d %s" % (K, tag))
        assert numpy.mean(it == ref["iters"]) >= 0.98
        handed = out["stats"]["docs_narrow"] + out["stats"]["docs_narrow_wide"] + out["stats"]["docs_long_compact"]
        print(tag, "K=%d" % K, "handed over", handed)
        if K > 32 and tag == "peaked":        # (with K <= 32 the register-tile kernel has no compact stage to hand over from)
            assert handed > 0


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_sweep(ctx, seed):
    """Randomised sweep over the number of topics (every compiled lane shape and owner width), corpus
    shape (incl. documents long enough for the streaming path) and asymmetric alpha."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    rs = numpy.random.RandomState(1000 + seed)
    K = int(rs.choice([1, 2, 5, 7, 11, 20, 31, 33, 48, 57, 80, 96, 97, 104, 105, 112, 130, 160, 256, 300]))
    V = int(rs.randint(max(40, K), 1500))
    D = int(rs.randint(3, 40))
    length = "zipf" if rs.rand() < 0.5 else "poisson"
    row_ptr, ids, cts = synthetic.synthetic_corpus(D, V, seed=seed, length=length, mean_len=int(rs.randint(5, 400)))
    eta = rs.gamma(rs.choice([0.05, 1.0, 100.0]), 1.0, (K, V)) + 1e-3
    alpha = rs.uniform(0.01, 1.5, K)
    ref = O.e_step(row_ptr, ids, cts, eta, alpha, 50, 1e-6, heldout=True, return_iters=True)
    ctx.set_corpus(0, row_ptr, ids, cts)
    out = ctx.estep(0, eta, alpha, 50, 1e-6, heldout=True)
    _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "seed=%d K=%d V=%d D=%d" % (seed, K, V, D))
    assert abs(out["words_ll"] - ref["words_ll"]) <= RTOL * abs(ref["words_ll"])
    it = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
    # trip counts: identical, except that one document sitting on the stop threshold may flip by one trip
    assert numpy.sum(it != ref["iters"]) <= 1 and numpy.max(numpy.abs(it - ref["iters"])) <= 1


def test_pathologically_long_document(ctx):
    """A document with more distinct terms than the streaming kernel holds in shared memory at a time (~9000) is
    walked in chunks, re-staged every trip; same results."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    K, V = 8, 16000
    n_long = 12000
    rs = numpy.random.RandomState(4)
    long_ids = numpy.sort(rs.choice(V, n_long, replace=False)).astype(numpy.int32)
    long_cts = rs.randint(1, 4, n_long).astype(numpy.int32)
    a = synthetic.synthetic_corpus(20, V, seed=3, length="poisson", mean_len=60)
    row_ptr = numpy.concatenate([a[0], [a[0][-1] + n_long]]).astype(numpy.int64)
    ids = numpy.concatenate([a[1], long_ids])
    cts = numpy.concatenate([a[2], long_cts])
    eta = synthetic.initial_eta(K, V, 0)
    alpha = numpy.full(K, 0.3)
    ref = O.e_step(row_ptr, ids, cts, eta, alpha, 50, 1e-6, return_iters=True)
    ctx.set_corpus(0, row_ptr, ids, cts)
    out = ctx.estep(0, eta, alpha, 50, 1e-6)
    _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "long document")
    assert out["stats"]["docs_streamed"] >= 1


def test_dead_topic_elimination_changes_nothing(ctx, monkeypatch):
    """PYLDA_COMPACT=0 (every trip at full width) and the default (topics with gamma_k == alpha_k dropped,
    trips continued on a 32-column compact tile) must agree to rounding, with identical trip counts and
    no revival; short, medium and long documents, symmetric alpha small enough for topics to die."""
    from pylda_b200 import synthetic
    K, V = 100, 20000
    row_ptr, ids, cts = synthetic.synthetic_corpus(3000, V, seed=41, length="zipf")
    eta = synthetic.initial_eta(K, V, 0)
    alpha = numpy.full(K, 1.0 / K)
    ctx.set_corpus(0, row_ptr, ids, cts)
    monkeypatch.setenv("PYLDA_COMPACT", "0")
    full = ctx.estep(0, eta, alpha, 50, 1e-6)
    it_full = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
    monkeypatch.delenv("PYLDA_COMPACT")
    ctx.estep(0, eta, alpha, 50, 1e-6)                       # (first use allocates the hand-over buffers)
    fast = ctx.estep(0, eta, alpha, 50, 1e-6)
    it_fast = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
    assert fast["stats"]["revived_docs"] == 0
    assert numpy.array_equal(it_full, it_fast)
    assert max_rel(fast["gamma"], full["gamma"]) <= 1e-12
    assert max_rel(fast["phi_ss"], full["phi_ss"], floor=PHI_FLOOR) <= 1e-10
    assert abs(fast["doc_ll"] - full["doc_ll"]) <= 1e-12 * abs(full["doc_ll"])
    assert fast["stats"]["kernel_ms"] < full["stats"]["kernel_ms"]


def test_narrow_stages_change_nothing(ctx, monkeypatch):
    """PYLDA_PARK=0 (every document finishes in the register-tile kernel) and the default (documents with at
    most 16 / 8 live topics finish in the narrow stages, estep_narrow.cuh) must agree to rounding: identical
    trip counts, gamma, ELBO -- and phi_ss down to its smallest entries: the statistics of the topics
    eliminated as dead (~1e-44 per entry at alpha = 0.01) are added through the per-word weight sums."""
    from oracle import estep_oracle as O
    from pylda_b200 import synthetic
    K, V = 100, 20000
    row_ptr, ids, cts = synthetic.synthetic_corpus(3000, V, seed=43, length="zipf")
    eta = synthetic.initial_eta(K, V, 0)
    alpha = numpy.full(K, 1.0 / K)
    ctx.set_corpus(0, row_ptr, ids, cts)
    monkeypatch.setenv("PYLDA_PARK", "0")
    base = ctx.estep(0, eta, alpha, 50, 1e-6)
    it_base = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
    assert base["stats"]["docs_narrow"] == 0 and base["stats"]["docs_narrow_wide"] == 0
    for park in ("16", "8"):
        monkeypatch.setenv("PYLDA_PARK", park)
        fast = ctx.estep(0, eta, alpha, 50, 1e-6)
        it_fast = ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"]
        st = fast["stats"]
        print("park", park, "narrow", st["docs_narrow_wide"], st["docs_narrow"], "kernel ms", st["kernel_ms"], base["stats"]["kernel_ms"])
        assert st["docs_narrow"] > 1000 and st["revived_docs"] == 0
        assert (st["docs_narrow_wide"] > 1000) == (park == "16")
        assert numpy.array_equal(it_base, it_fast)
        assert max_rel(fast["gamma"], base["gamma"]) <= 1e-11
        assert max_rel(fast["phi_ss"], base["phi_ss"], floor=1e-250) <= 1e-9      # relative, tiny entries included
        assert abs(fast["doc_ll"] - base["doc_ll"]) <= 1e-12 * abs(base["doc_ll"])
    # ... and against the oracle on a sub-corpus, again down to the tiny entries
    sub = 400
    rp, nz = row_ptr[:sub + 1], int(row_ptr[sub])
    ref = O.e_step(rp, ids[:nz], cts[:nz], eta, alpha, 50, 1e-6, return_iters=True)
    ctx.set_corpus(0, rp, ids[:nz], cts[:nz])
    monkeypatch.delenv("PYLDA_PARK")
    out = ctx.estep(0, eta, alpha, 50, 1e-6)
    _check(out, ref["gamma"], ref["phi_ss"], ref["doc_ll"], "narrow vs oracle")
    assert out["stats"]["docs_narrow"] > 100
    assert max_rel(out["phi_ss"], ref["phi_ss"], floor=1e-250) <= RTOL
    assert numpy.array_equal(ctx.get_results(0, gamma=False, phi=False, iters=True)["iters"], ref["iters"])


def test_narrow_stages_off_when_alpha_is_large(ctx):
    """The hand-over relies on exp(psi(alpha_k)) being far below ulp(alpha_k): with alpha = 0.05 the safety
    bound disables it (topics then never die bit for bit either)."""
    from pylda_b200 import synthetic
    K, V = 100, 5000
    row_ptr, ids, cts = synthetic.synthetic_corpus(200, V, seed=44, length="zipf")
    ctx.set_corpus(0, row_ptr, ids, cts)
    out = ctx.estep(0, synthetic.initial_eta(K, V, 0), numpy.full(K, 0.05), 50, 1e-6)
    assert out["stats"]["docs_narrow"] == 0 and out["stats"]["docs_narrow_wide"] == 0


def test_full_width_redo_when_an_eliminated_topic_comes_back(ctx, monkeypatch):
    """The kernels verify per document that every topic eliminated as dead stays dead (stats.revived_docs).
    If that ever fails the library redoes the E-step with the elimination and the narrow stages off; the
    test hook forces that path and the results must be the plain full-width ones."""
    g = load_golden("zipf48_k100")
    ctx.set_corpus(0, g["row_ptr"], g["ids"], g["cts"])
    plain = ctx.estep(0, g["eta"], g["alpha"], 50, 1e-6)
    assert plain["stats"]["revived_docs"] == 0 and plain["stats"]["docs_narrow"] > 0
    monkeypatch.setenv("PYLDA_TEST_FORCE_REDO", "1")
    redo = ctx.estep(0, g["eta"], g["alpha"], 50, 1e-6)
    monkeypatch.delenv("PYLDA_TEST_FORCE_REDO")
    assert redo["stats"]["revived_docs"] == 1 and redo["stats"]["docs_narrow"] == 0
    _check(dict(redo, stats=None), g["gamma"], g["phi_ss"], g["doc_ll"], "redo")
    assert max_rel(redo["gamma"], plain["gamma"]) <= 1e-11
