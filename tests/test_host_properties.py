"""Property tests (hypothesis) of the integer host logic on the path: CSR packing of the reference's
parsed-corpus type (variational_bayes.py:98-130) and the nnz-balanced document sharding.
Integer work: everything must be bit-exact."""
import numpy
from hypothesis import given, settings, strategies as st

from pylda_b200 import native
from pylda_b200.variational_bayes import pack_parsed_corpus

doc = st.lists(st.tuples(st.integers(0, 5000), st.integers(1, 40)), min_size=1, max_size=30, unique_by=lambda t: t[0])
corpus = st.lists(doc, min_size=0, max_size=40)


def _parsed(docs):
    word_ids = [numpy.array([w for w, _ in d]) for d in docs]
    word_cts = [numpy.array([c for _, c in d])[numpy.newaxis, :] for d in docs]     # (1, n_d) as the reference
    return word_ids, word_cts


@settings(max_examples=60, deadline=None)
@given(corpus)
def test_pack_parsed_corpus_is_lossless(docs):
    word_ids, word_cts = _parsed(docs)
    row_ptr, ids, cts = pack_parsed_corpus((word_ids, word_cts))
    assert row_ptr.dtype == numpy.int64 and ids.dtype == numpy.int32 and cts.dtype == numpy.int32
    assert row_ptr[0] == 0 and len(row_ptr) == len(docs) + 1 and row_ptr[-1] == len(ids) == len(cts)
    for d, (w, c) in enumerate(zip(word_ids, word_cts)):
        a, b = int(row_ptr[d]), int(row_ptr[d + 1])
        assert numpy.array_equal(ids[a:b], w) and numpy.array_equal(cts[a:b], c.reshape(-1))


@settings(max_examples=60, deadline=None)
@given(corpus, st.integers(1, 9))
def test_shards_partition_the_corpus(docs, n_ranks):
    row_ptr, ids, cts = pack_parsed_corpus(_parsed(docs))
    b = native.shard_bounds(row_ptr, n_ranks)
    assert len(b) == n_ranks + 1 and b[0] == 0 and b[-1] == len(docs) and numpy.all(numpy.diff(b) >= 0)
    parts = [native.shard_csr(row_ptr, ids, cts, int(b[r]), int(b[r + 1])) for r in range(n_ranks)]
    assert numpy.array_equal(numpy.concatenate([p[1] for p in parts]) if parts else ids, ids)
    assert numpy.array_equal(numpy.concatenate([p[2] for p in parts]) if parts else cts, cts)
    assert sum(len(p[0]) - 1 for p in parts) == len(docs)
    for p in parts:
        assert p[0][0] == 0 and p[0][-1] == len(p[1]) and numpy.all(numpy.diff(p[0]) >= 1 if len(p[0]) > 1 else True)
    if len(docs):
        longest = int(numpy.diff(row_ptr).max())
        nnz = numpy.array([len(p[1]) for p in parts])
        assert nnz.max() - nnz.min() <= 2 * longest + 1 or n_ranks > len(docs)


def test_csr_cache_round_trip_and_shards(tmp_path):
    """On-disk CSR cache: what rank 0 writes is bit-identical to the parsed corpus, every rank's
    shard is exactly its nnz-balanced slice, a different vocabulary order gives a different key, and the lazy
    stand-in for (word_ids, word_cts) materialises the reference's list type."""
    import numpy
    from pylda_b200 import corpus_cache, native
    from pylda_b200 import variational_bayes as vb
    V = 300
    rs = numpy.random.RandomState(3)
    docs = [" ".join("w%d" % t for t in rs.randint(0, V + 20, rs.randint(1, 60))) for _ in range(400)] + ["zzz"]
    vocab = ["w%d" % i for i in range(V)]
    a = vb.VariationalBayes(); a.parse_vocabulary(vocab)
    parsed, csr = a._parse(docs)
    csr = csr if csr is not None else vb.pack_parsed_corpus(parsed)
    key, vs, cs = corpus_cache.cache_key(docs, a._index_to_type)
    assert corpus_cache.open_entry(str(tmp_path), key) is None
    corpus_cache.save(str(tmp_path), key, csr, len(docs) - len(parsed[0]), vs, cs)
    entry = corpus_cache.open_entry(str(tmp_path), key)
    meta, row_ptr, ids, cts = entry
    D = len(parsed[0])
    assert meta["D"] == D and 390 <= D <= 400 and meta["dropped"] == len(docs) - D >= 1
    assert numpy.array_equal(row_ptr, csr[0]) and numpy.array_equal(ids, csr[1]) and numpy.array_equal(cts, csr[2])
    bounds = native.shard_bounds(csr[0], 3)
    seen = 0
    for r in range(3):
        lo, hi, (rp, i2, c2) = corpus_cache.load_shard(entry, r, 3)
        assert (lo, hi) == (int(bounds[r]), int(bounds[r + 1])) and rp[0] == 0
        a0, a1 = int(csr[0][lo]), int(csr[0][hi])
        assert numpy.array_equal(rp, csr[0][lo:hi + 1] - a0) and numpy.array_equal(i2, csr[1][a0:a1]) and numpy.array_equal(c2, csr[2][a0:a1])
        seen += hi - lo
    assert seen == D
    lazy = corpus_cache.LazyParsed(entry)
    assert lazy.number_of_documents == D and len(lazy) == 2
    for x, y in zip(lazy[0], parsed[0]):
        assert numpy.array_equal(x, y)
    for x, y in zip(lazy[1], parsed[1]):
        assert x.shape == y.shape and numpy.array_equal(x, y)
    b = vb.VariationalBayes(); b.parse_vocabulary(list(reversed(vocab)))
    assert corpus_cache.cache_key(docs, b._index_to_type)[0] != key
