"""On-disk CSR form of a parsed corpus (reference variational_bayes.py:98-130 parse_data,
launch_train.py:101-116).

The reference parses the text corpus in every process that needs it.  At the multi-GPU configurations (one process
per GPU, 1M-10M documents) that is 8x redundant host work and 8 full copies of the corpus in host memory.  With a
cache directory (environment PYLDA_CSR_CACHE, or launch_train --csr_cache) rank 0 parses ONCE (native parser when
the text is ASCII) and writes

    <dir>/<key>/row_ptr.npy  int64 (D+1)     <dir>/<key>/ids.npy  int32 (nnz)     <dir>/<key>/cts.npy  int32 (nnz)
    <dir>/<key>/meta.json    {"D", "nnz", "dropped", "vocab_sha1", "corpus_sha1"}      (written last: the commit mark)

and every rank memory-maps the three arrays and reads only the rows of its own nnz-balanced shard.  The key is a
digest of the vocabulary in type-id order and of the corpus text, so a cache can never be applied to another
vocabulary ordering (type ids come from set() iteration order, inferencer.py:63-65) or another corpus.
Integer data only: the token indexing round-trips bit-exactly (tests/test_host_properties.py).
"""
import hashlib
import json
import os
import time

import numpy


def cache_key(corpus, index_to_type):
    h = hashlib.sha1()
    for i in range(len(index_to_type)):
        h.update(index_to_type[i].encode("utf-8", "surrogatepass"))
        h.update(b"\n")
    vocab_sha1 = h.hexdigest()
    h = hashlib.sha1()
    h.update(str(len(corpus)).encode())
    for line in corpus:
        h.update(line.encode("utf-8", "surrogatepass"))
        h.update(b"\n")
    corpus_sha1 = h.hexdigest()
    return vocab_sha1[:12] + "-" + corpus_sha1[:12], vocab_sha1, corpus_sha1


def save(directory, key, csr, dropped, vocab_sha1, corpus_sha1):
    path = os.path.join(directory, key)
    os.makedirs(path, exist_ok=True)
    row_ptr, ids, cts = csr
    for name, arr, dt in (("row_ptr", row_ptr, numpy.int64), ("ids", ids, numpy.int32), ("cts", cts, numpy.int32)):
        tmp = os.path.join(path, name + ".tmp.npy")
        numpy.save(tmp, numpy.ascontiguousarray(arr, dtype=dt))
        os.replace(tmp, os.path.join(path, name + ".npy"))
    meta = dict(D=int(len(row_ptr) - 1), nnz=int(len(ids)), dropped=int(dropped), vocab_sha1=vocab_sha1, corpus_sha1=corpus_sha1)
    tmp = os.path.join(path, "meta.json.tmp")
    with open(tmp, "w") as f:
        json.dump(meta, f)
    os.replace(tmp, os.path.join(path, "meta.json"))
    return path


def wait_for(directory, key, timeout=3600.0):
    """Block until rank 0 has committed the cache entry; returns its meta."""
    meta_path = os.path.join(directory, key, "meta.json")
    deadline = time.time() + timeout
    while not os.path.exists(meta_path):
        if time.time() > deadline:
            raise RuntimeError("pylda_b200: timed out waiting for the CSR cache entry %s" % meta_path)
        time.sleep(0.2)
    with open(meta_path) as f:
        return json.load(f)


def open_entry(directory, key):
    """(meta, row_ptr, ids, cts) with the arrays memory-mapped read-only, or None when the entry does not exist."""
    path = os.path.join(directory, key)
    meta_path = os.path.join(path, "meta.json")
    if not os.path.exists(meta_path):
        return None
    with open(meta_path) as f:
        meta = json.load(f)
    arrs = [numpy.load(os.path.join(path, n + ".npy"), mmap_mode="r") for n in ("row_ptr", "ids", "cts")]
    if arrs[0].shape[0] != meta["D"] + 1 or arrs[1].shape[0] != meta["nnz"] or arrs[2].shape[0] != meta["nnz"]:
        return None
    return (meta,) + tuple(arrs)


def load_shard(entry, rank, world):
    """(lo, hi, (row_ptr, ids, cts)) of this rank's nnz-balanced contiguous shard, copied out of the mapped files
    (only these rows are read from disk)."""
    from . import native
    _, row_ptr, ids, cts = entry
    D = row_ptr.shape[0] - 1
    if world <= 1:
        lo, hi = 0, D
    else:
        b = native.shard_bounds(numpy.asarray(row_ptr), world)
        lo, hi = int(b[rank]), int(b[rank + 1])
    a, e = int(row_ptr[lo]), int(row_ptr[hi])
    return lo, hi, (numpy.asarray(row_ptr[lo:hi + 1], dtype=numpy.int64) - a, numpy.array(ids[a:e]), numpy.array(cts[a:e]))


class LazyParsed(object):
    """Stands in for the reference's parsed-corpus tuple (word_ids, word_cts) (variational_bayes.py:127-130) when the
    corpus lives in a CSR cache: the two lists of per-document arrays are only built if somebody indexes them."""

    def __init__(self, entry):
        self._entry = entry
        self._lists = None

    def _materialize(self):
        if self._lists is None:
            _, row_ptr, ids, cts = self._entry
            bounds = numpy.asarray(row_ptr[1:-1])
            word_ids = numpy.split(numpy.asarray(ids, dtype=numpy.int64), bounds) if row_ptr.shape[0] > 1 else []
            word_cts = [c[numpy.newaxis, :] for c in numpy.split(numpy.asarray(cts, dtype=numpy.int64), bounds)] \
                if row_ptr.shape[0] > 1 else []
            self._lists = (word_ids, word_cts)
        return self._lists

    @property
    def number_of_documents(self):
        return int(self._entry[1].shape[0] - 1)

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return self._materialize()[i]

    def __iter__(self):
        return iter(self._materialize())

    def __getstate__(self):          # pickled models carry the plain tuple, like the reference's
        return {"lists": self._materialize()}

    def __setstate__(self, state):
        self._entry = None
        self._lists = state["lists"]
