"""ctypes binding of include/pylda_b200.h -- the stub a PyLDA maintainer would add inside
VariationalBayes.e_step (variational_bayes.py:132).  No torch, no CPU fallback: if the shared
library or a B200 is missing, construction raises.
"""
import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpylda_b200.so")
ABI_VERSION = 10
NCCL_ID_BYTES = 128

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)


class Stats(ctypes.Structure):
    _fields_ = [
        ("n_docs", ctypes.c_int64),
        ("nnz", ctypes.c_int64),
        ("inner_iters", ctypes.c_int64),
        ("docs_at_cap", ctypes.c_int64),
        ("prep_ms", ctypes.c_double),
        ("kernel_ms", ctypes.c_double),
        ("post_ms", ctypes.c_double),
        ("total_ms", ctypes.c_double),
        ("algo_read_bytes", ctypes.c_double),
        ("algo_total_bytes", ctypes.c_double),
        ("n_launches", ctypes.c_int32),
        ("n_estep_launches", ctypes.c_int32),
        ("docs_resident", ctypes.c_int64),
        ("docs_streamed", ctypes.c_int64),
        ("row_trips", ctypes.c_double),
        ("revived_docs", ctypes.c_int64),
        ("docs_narrow_wide", ctypes.c_int64),
        ("docs_narrow", ctypes.c_int64),
        ("allreduce_ms", ctypes.c_double),
        ("gamma_rows_early", ctypes.c_int64),
        ("docs_long_compact", ctypes.c_int64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# every symbol include/pylda_b200.h declares: name -> (restype, argtypes)
_SIGNATURES = {
    "pylda_abi_version": (ctypes.c_int, []),
    "pylda_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "pylda_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pylda_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "pylda_set_corpus": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                        _c_int64_p, _c_int32_p, _c_int32_p]),
    "pylda_get_corpus": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int64_p, _c_int32_p, _c_int32_p]),
    "pylda_estep": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   _c_double_p, _c_double_p, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                   _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                   ctypes.POINTER(Stats)]),
    "pylda_set_model": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p]),
    "pylda_estep_resident": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                            ctypes.c_int, ctypes.c_int, ctypes.POINTER(Stats)]),
    "pylda_get_results": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p,
                                         _c_double_p, _c_double_p, _c_int32_p]),
    "pylda_mstep_resident": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, _c_double_p, _c_double_p]),
    "pylda_get_eta": (ctypes.c_int, [ctypes.c_void_p, _c_double_p]),
    "pylda_set_alpha": (ctypes.c_int, [ctypes.c_void_p, _c_double_p]),
    "pylda_dirichlet_expectation": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                   _c_double_p, _c_double_p]),
    "pylda_top_words": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int32_p, _c_double_p]),
    "pylda_special": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, _c_double_p, _c_double_p]),
    "pylda_comm_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "pylda_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]),
    "pylda_comm_allreduce_sum": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, ctypes.c_int64]),
    "pylda_host_register": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    "pylda_host_unregister": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pylda_parse_corpus": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_void_p)]),
    "pylda_parsed_dims": (ctypes.c_int, [ctypes.c_void_p, _c_int64_p, _c_int64_p, _c_int64_p]),
    "pylda_parsed_copy": (ctypes.c_int, [ctypes.c_void_p, _c_int64_p, _c_int32_p, _c_int32_p]),
    "pylda_parsed_free": (ctypes.c_int, [ctypes.c_void_p]),
    "pylda_device_name": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]),
    "pylda_sm_count": (ctypes.c_int, [ctypes.c_void_p]),
}

_lib = None


def load_library():
    """dlopen the in-tree library and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("pylda_b200: %s is missing -- build it with `python -m pylda_b200.build` "
                           "(there is no CPU fallback for the E-step)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.pylda_abi_version() != ABI_VERSION:
        raise RuntimeError("pylda_b200: ABI mismatch (library %d, binding %d)" % (lib.pylda_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(_c_double_p)


def _checked(a, dtype, shape, name):
    a = numpy.ascontiguousarray(a, dtype=dtype)
    if tuple(a.shape) != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), a.shape))
    return a


class EStepContext(object):
    """One device context (one process per GPU).  Owns nothing on the host side except the
    opaque handle; every numpy array passed in stays caller-owned."""

    def __init__(self, device=0):
        self._h = None
        lib = load_library()
        h = ctypes.c_void_p()
        if lib.pylda_create(ctypes.byref(h), int(device)) != 0:
            raise RuntimeError("pylda_create failed: %s" % lib.pylda_last_error(None).decode())
        self._lib = lib
        self._h = h
        self._corpus_dims = {}
        self.K = self.V = None
        self.last_stats = None

    # -- lifecycle --------------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            self._lib.pylda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %s" % (what, self._lib.pylda_last_error(self._h).decode()))

    # -- corpus -----------------------------------------------------------------------------
    def set_corpus(self, slot, row_ptr, ids, cts):
        row_ptr = numpy.ascontiguousarray(row_ptr, dtype=numpy.int64)
        ids = numpy.ascontiguousarray(ids, dtype=numpy.int32)
        cts = numpy.ascontiguousarray(cts, dtype=numpy.int32)
        D = row_ptr.shape[0] - 1
        nnz = ids.shape[0]
        if cts.shape[0] != nnz:
            raise ValueError("ids and cts must have the same length")
        self._check(self._lib.pylda_set_corpus(self._h, slot, D, nnz, row_ptr.ctypes.data_as(_c_int64_p),
                                               ids.ctypes.data_as(_c_int32_p), cts.ctypes.data_as(_c_int32_p)),
                    "pylda_set_corpus")
        self._corpus_dims[slot] = (D, nnz)

    def get_corpus(self, slot):
        D, nnz = self._corpus_dims[slot]
        row_ptr = numpy.empty(D + 1, dtype=numpy.int64)
        ids = numpy.empty(nnz, dtype=numpy.int32)
        cts = numpy.empty(nnz, dtype=numpy.int32)
        self._check(self._lib.pylda_get_corpus(self._h, slot, row_ptr.ctypes.data_as(_c_int64_p),
                                               ids.ctypes.data_as(_c_int32_p), cts.ctypes.data_as(_c_int32_p)),
                    "pylda_get_corpus")
        return row_ptr, ids, cts

    # -- the reference call: e_step with host buffers ----------------------------------------
    def estep(self, slot, eta, alpha, max_iter=50, tol=1e-6, heldout=False,
              want_gamma=True, want_phi=True, want_alpha_ss=False, gamma_out=None, phi_out=None):
        eta = numpy.ascontiguousarray(eta, dtype=numpy.float64)
        K, V = eta.shape
        alpha = _checked(alpha, numpy.float64, (K,), "alpha")
        D, _ = self._corpus_dims[slot]
        gamma = phi = None
        if want_gamma:
            gamma = numpy.empty((D, K), dtype=numpy.float64) if gamma_out is None else gamma_out
            assert gamma.shape == (D, K) and gamma.dtype == numpy.float64 and gamma.flags.c_contiguous
        if want_phi:
            phi = numpy.empty((K, V), dtype=numpy.float64) if phi_out is None else phi_out
            assert phi.shape == (K, V) and phi.dtype == numpy.float64 and phi.flags.c_contiguous
        ass = numpy.empty(K, dtype=numpy.float64) if want_alpha_ss else None
        doc_ll = ctypes.c_double(0.0)
        words_ll = ctypes.c_double(0.0)
        st = Stats()
        self._check(self._lib.pylda_estep(self._h, slot, K, V, _dp(eta), _dp(alpha), int(max_iter), float(tol),
                                          1 if heldout else 0, _dp(gamma), _dp(phi), _dp(ass),
                                          ctypes.byref(doc_ll), ctypes.byref(words_ll), ctypes.byref(st)),
                    "pylda_estep")
        self.K, self.V = K, V
        self.last_stats = st.as_dict()
        return dict(gamma=gamma, phi_ss=phi, alpha_ss=ass, doc_ll=doc_ll.value, words_ll=words_ll.value,
                    stats=self.last_stats)

    # -- split form: inputs resident in HBM ---------------------------------------------------
    def set_model(self, eta, alpha):
        eta = numpy.ascontiguousarray(eta, dtype=numpy.float64)
        K, V = eta.shape
        alpha = _checked(alpha, numpy.float64, (K,), "alpha")
        self._check(self._lib.pylda_set_model(self._h, K, V, _dp(eta), _dp(alpha)), "pylda_set_model")
        self.K, self.V = K, V

    def set_alpha(self, alpha):
        alpha = _checked(alpha, numpy.float64, (self.K,), "alpha")
        self._check(self._lib.pylda_set_alpha(self._h, _dp(alpha)), "pylda_set_alpha")

    def estep_resident(self, slot, max_iter=50, tol=1e-6, heldout=False, want_alpha_ss=False):
        st = Stats()
        self._check(self._lib.pylda_estep_resident(self._h, slot, int(max_iter), float(tol), 1 if heldout else 0,
                                                   1 if want_alpha_ss else 0, ctypes.byref(st)),
                    "pylda_estep_resident")
        self.last_stats = st.as_dict()
        return self.last_stats

    def get_results(self, slot, gamma=True, phi=True, alpha_ss=False, iters=False):
        D, _ = self._corpus_dims[slot]
        K, V = self.K, self.V
        g = numpy.empty((D, K), dtype=numpy.float64) if gamma else None
        p = numpy.empty((K, V), dtype=numpy.float64) if phi else None
        a = numpy.empty(K, dtype=numpy.float64) if alpha_ss else None
        it = numpy.empty(D, dtype=numpy.int32) if iters else None
        doc_ll = ctypes.c_double(0.0)
        words_ll = ctypes.c_double(0.0)
        self._check(self._lib.pylda_get_results(self._h, slot, _dp(g), _dp(p), _dp(a), ctypes.byref(doc_ll),
                                                ctypes.byref(words_ll),
                                                None if it is None else it.ctypes.data_as(_c_int32_p)),
                    "pylda_get_results")
        return dict(gamma=g, phi_ss=p, alpha_ss=a, iters=it, doc_ll=doc_ll.value, words_ll=words_ll.value)

    def mstep_resident(self, alpha_beta, want_eta=True):
        eta = numpy.empty((self.K, self.V), dtype=numpy.float64) if want_eta else None
        t = ctypes.c_double(0.0)
        self._check(self._lib.pylda_mstep_resident(self._h, float(alpha_beta), ctypes.byref(t), _dp(eta)),
                    "pylda_mstep_resident")
        return t.value, eta

    def get_eta(self):
        eta = numpy.empty((self.K, self.V), dtype=numpy.float64)
        self._check(self._lib.pylda_get_eta(self._h, _dp(eta)), "pylda_get_eta")
        return eta

    def top_words(self, top):
        """(idx (K, top) int32, prob (K, top) float64): the `top` most probable words of every topic of the model on
        the device, most probable first (device side of export_beta, variational_bayes.py:326-341)."""
        idx = numpy.empty((self.K, top), dtype=numpy.int32)
        prob = numpy.empty((self.K, top), dtype=numpy.float64)
        self._check(self._lib.pylda_top_words(self._h, int(top), idx.ctypes.data_as(_c_int32_p), _dp(prob)), "pylda_top_words")
        return idx, prob

    # -- helpers exposed for parity tests -----------------------------------------------------
    def dirichlet_expectation(self, eta):
        eta = numpy.ascontiguousarray(eta, dtype=numpy.float64)
        K, V = eta.shape
        out = numpy.empty((K, V), dtype=numpy.float64)
        self._check(self._lib.pylda_dirichlet_expectation(self._h, K, V, _dp(eta), _dp(out)),
                    "pylda_dirichlet_expectation")
        return out

    def special(self, which, x):
        x = numpy.ascontiguousarray(x, dtype=numpy.float64).reshape(-1)
        out = numpy.empty_like(x)
        code = {"digamma": 0, "exp_digamma": 1, "lgamma": 2, "rcp": 3, "exp_digamma_v2": 4}[which]
        self._check(self._lib.pylda_special(self._h, code, x.shape[0], _dp(x), _dp(out)), "pylda_special")
        return out

    # -- multi-GPU ---------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        lib = load_library()
        buf = ctypes.create_string_buffer(NCCL_ID_BYTES)
        if lib.pylda_comm_unique_id(buf) != 0:
            raise RuntimeError("pylda_comm_unique_id failed: %s" % lib.pylda_last_error(None).decode())
        return buf.raw

    def comm_init(self, n_ranks, rank, unique_id):
        assert len(unique_id) == NCCL_ID_BYTES
        self._check(self._lib.pylda_comm_init(self._h, int(n_ranks), int(rank), unique_id), "pylda_comm_init")

    def allreduce_sum(self, array):
        """In-place sum over ranks of a C-contiguous float64 array (no-op on a single rank)."""
        assert array.dtype == numpy.float64 and array.flags.c_contiguous
        self._check(self._lib.pylda_comm_allreduce_sum(self._h, _dp(array), array.size), "pylda_comm_allreduce_sum")
        return array

    def pin(self, array):
        """Page-lock a numpy array in place (cudaHostRegister)."""
        self._check(self._lib.pylda_host_register(self._h, ctypes.c_void_p(array.ctypes.data), array.nbytes),
                    "pylda_host_register")

    def unpin(self, array):
        self._check(self._lib.pylda_host_unregister(self._h, ctypes.c_void_p(array.ctypes.data)),
                    "pylda_host_unregister")

    def device_name(self):
        buf = ctypes.create_string_buffer(256)
        self._lib.pylda_device_name(self._h, buf, 256)
        return buf.value.decode()

    def sm_count(self):
        return self._lib.pylda_sm_count(self._h)


def parse_corpus(corpus, index_to_type, n_threads=0):
    """Native text -> CSR (pylda_parse_corpus): corpus = iterable of document strings, index_to_type =
    {type id: word}.  Returns (row_ptr, ids, cts, dropped) or None when the text is not pure ASCII
    (str.split() then also splits on Unicode spaces, which only the Python path reproduces).  Host code,
    no device needed."""
    lib = load_library()
    try:
        text = ("\n".join(corpus) + "\n").encode("ascii") if len(corpus) else b""
        vocab = "\n".join(index_to_type[i] for i in range(len(index_to_type))).encode("ascii")
    except UnicodeEncodeError:
        return None
    if b"\n" in vocab and len(vocab.split(b"\n")) != len(index_to_type):
        return None
    if text.count(b"\n") != len(corpus):          # a document with an embedded newline: leave it to Python
        return None
    h = ctypes.c_void_p()
    if lib.pylda_parse_corpus(text, len(text), vocab, len(vocab), int(n_threads), ctypes.byref(h)) != 0:
        raise RuntimeError("pylda_parse_corpus failed")
    try:
        D, nnz, dropped = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        lib.pylda_parsed_dims(h, ctypes.byref(D), ctypes.byref(nnz), ctypes.byref(dropped))
        row_ptr = numpy.empty(D.value + 1, dtype=numpy.int64)
        ids = numpy.empty(nnz.value, dtype=numpy.int32)
        cts = numpy.empty(nnz.value, dtype=numpy.int32)
        lib.pylda_parsed_copy(h, row_ptr.ctypes.data_as(_c_int64_p), ids.ctypes.data_as(_c_int32_p),
                              cts.ctypes.data_as(_c_int32_p))
    finally:
        lib.pylda_parsed_free(h)
    return row_ptr, ids, cts, int(dropped.value)


def shard_bounds(row_ptr, n_ranks):
    """Contiguous document ranges balanced by nnz: returns n_ranks+1
    document indices.  Pure integer host logic, shared by every rank."""
    row_ptr = numpy.asarray(row_ptr, dtype=numpy.int64)
    D = row_ptr.shape[0] - 1
    nnz = int(row_ptr[-1])
    targets = (numpy.arange(1, n_ranks, dtype=numpy.float64) * nnz / n_ranks)
    cuts = numpy.searchsorted(row_ptr, targets, side="left")
    bounds = numpy.concatenate([[0], numpy.minimum(cuts, D), [D]]).astype(numpy.int64)
    return numpy.maximum.accumulate(bounds)


def shard_csr(row_ptr, ids, cts, lo, hi):
    """CSR slice of documents [lo, hi) with a re-based row_ptr."""
    a, b = int(row_ptr[lo]), int(row_ptr[hi])
    return (numpy.asarray(row_ptr[lo:hi + 1], dtype=numpy.int64) - a, ids[a:b], cts[a:b])
