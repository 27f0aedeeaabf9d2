"""Seeded synthetic bag-of-words corpora in CSR form (BASELINE.json configs 2/3/5).

Not reference code: the reference ships no generator.  Word distribution
p_v ~ 1/rank^word_exponent with ranks a random permutation of 1..V (hot words are not
adjacent in memory); document length N_d ~ max(1, Poisson(mean_len)) ("poisson") or
clip(floor(Zipf(2.2) * 100 / 3.75), 8, 4096) ("zipf").  Per document: draw N_d tokens,
unique -> term ids ascending with counts -- the same (ids, counts) content that
variational_bayes.py:98-130 (parse_data) produces for the rendered text.
"""
import numpy


def synthetic_corpus(D, V, seed, length="poisson", mean_len=100, word_exponent=1.0, chunk=1 << 18):
    rng = numpy.random.default_rng(seed)
    ranks = rng.permutation(V) + 1
    p = 1.0 / ranks.astype(numpy.float64) ** word_exponent
    cdf = numpy.cumsum(p / p.sum())
    cdf[-1] = 1.0
    if length == "poisson":
        N = numpy.maximum(1, rng.poisson(mean_len, size=D)).astype(numpy.int64)
    elif length == "zipf":
        N = numpy.clip(numpy.floor(rng.zipf(2.2, size=D) * (100.0 / 3.75)), 8, 4096).astype(numpy.int64)
    else:
        raise ValueError(length)
    row_ptr = numpy.zeros(D + 1, dtype=numpy.int64)
    ids_parts, cts_parts = [], []
    for lo in range(0, D, chunk):
        hi = min(D, lo + chunk)
        n = N[lo:hi]
        total = int(n.sum())
        doc_of = numpy.repeat(numpy.arange(hi - lo, dtype=numpy.int64), n)
        toks = numpy.searchsorted(cdf, rng.random(total), side="left").astype(numpy.int64)
        numpy.minimum(toks, V - 1, out=toks)
        key = doc_of * V + toks
        key.sort()
        keep = numpy.empty(total, dtype=bool)
        keep[0] = True
        numpy.not_equal(key[1:], key[:-1], out=keep[1:])
        starts = numpy.flatnonzero(keep)
        uniq = key[starts]
        counts = numpy.diff(numpy.append(starts, total))
        udoc = uniq // V
        ids_parts.append((uniq - udoc * V).astype(numpy.int32))
        cts_parts.append(counts.astype(numpy.int32))
        row_ptr[lo + 1:hi + 1] = numpy.bincount(udoc, minlength=hi - lo)
    numpy.cumsum(row_ptr, out=row_ptr)
    return row_ptr, numpy.concatenate(ids_parts), numpy.concatenate(cts_parts)


def render_text(row_ptr, ids, cts):
    """CSR -> the reference's input format: one whitespace-tokenised document per line,
    token "w<id>" repeated count times (vocabulary is then w0..w{V-1})."""
    docs = []
    for d in range(len(row_ptr) - 1):
        a, b = int(row_ptr[d]), int(row_ptr[d + 1])
        toks = []
        for i, c in zip(ids[a:b], cts[a:b]):
            toks.extend(["w%d" % i] * int(c))
        docs.append(" ".join(toks))
    return docs


def initial_eta(K, V, seed=0):
    """variational_bayes.py:95 -- eta0 ~ Gamma(100, 1/100), drawn from numpy's legacy
    stream so numpy.random.seed(seed); numpy.random.gamma(...) gives the same matrix."""
    return numpy.random.RandomState(seed).gamma(100., 1. / 100., (K, V))


def lda_corpus(D, V, seed, topics=50, topic_concentration=0.05, doc_concentration=0.2, chunk=1 << 17):
    """Seeded corpus drawn from an actual LDA model (bench.py `warm_lda`): `topics` word distributions
    phi_k ~ Dirichlet(topic_concentration) over V types, per document theta_d ~ Dirichlet(doc_concentration),
    Zipf document lengths as in synthetic_corpus(length="zipf"), every token z ~ theta_d, w ~ phi_z.
    Unlike the single-Zipf generator this corpus has topic structure to converge to, so after a few EM
    iterations the per-document fixed point stops after ~10 trips instead of running to the cap."""
    rng = numpy.random.default_rng(seed)
    phi = rng.gamma(topic_concentration, 1.0, size=(topics, V))
    phi /= phi.sum(axis=1, keepdims=True)
    cdf = numpy.cumsum(phi, axis=1)
    cdf[:, -1] = 1.0
    N = numpy.clip(numpy.floor(rng.zipf(2.2, size=D) * (100.0 / 3.75)), 8, 4096).astype(numpy.int64)
    row_ptr = numpy.zeros(D + 1, dtype=numpy.int64)
    ids_parts, cts_parts = [], []
    for lo in range(0, D, chunk):
        hi = min(D, lo + chunk)
        n = N[lo:hi]
        theta = rng.gamma(doc_concentration, 1.0, size=(hi - lo, topics)) + 1e-12
        theta /= theta.sum(axis=1, keepdims=True)
        per_topic = rng.multinomial(n, theta)                       # (docs, topics) token counts
        keys = []
        for k in range(topics):
            c = per_topic[:, k]
            tot = int(c.sum())
            if tot == 0:
                continue
            doc_of = numpy.repeat(numpy.arange(hi - lo, dtype=numpy.int64), c)
            toks = numpy.searchsorted(cdf[k], rng.random(tot), side="left").astype(numpy.int64)
            numpy.minimum(toks, V - 1, out=toks)
            keys.append(doc_of * V + toks)
        key = numpy.concatenate(keys)
        key.sort()
        keep = numpy.empty(key.shape[0], dtype=bool)
        keep[0] = True
        numpy.not_equal(key[1:], key[:-1], out=keep[1:])
        starts = numpy.flatnonzero(keep)
        uniq = key[starts]
        counts = numpy.diff(numpy.append(starts, key.shape[0]))
        udoc = uniq // V
        ids_parts.append((uniq - udoc * V).astype(numpy.int32))
        cts_parts.append(counts.astype(numpy.int32))
        row_ptr[lo + 1:hi + 1] = numpy.bincount(udoc, minlength=hi - lo)
    numpy.cumsum(row_ptr, out=row_ptr)
    return row_ptr, numpy.concatenate(ids_parts), numpy.concatenate(cts_parts)
