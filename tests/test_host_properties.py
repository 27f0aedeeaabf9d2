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
