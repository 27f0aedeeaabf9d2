"""Python-3 mirror of /root/reference/inferencer.py (the abstract Inferencer API, :29-87).

Same class name, method names, argument meaning and attributes, so that callers written
against the reference (launch_train.py:187-204, launch_test.py:92-95, hybrid.py:23) keep
working.  Nothing here is on the hot path.
"""
import numpy
import scipy.special


def compute_dirichlet_expectation(dirichlet_parameter):
    """E[log theta] under Dirichlet(dirichlet_parameter): psi(x) - psi(sum x), along the last
    axis of a 1-D or 2-D array (inferencer.py:15-18).  Host version, used by export_beta; the
    E-step uses the device producer (csrc/prep_kernels.cuh)."""
    x = numpy.asarray(dirichlet_parameter)
    if x.ndim == 1:
        return scipy.special.psi(x) - scipy.special.psi(x.sum())
    return scipy.special.psi(x) - scipy.special.psi(x.sum(axis=1))[:, numpy.newaxis]


def parse_vocabulary(vocab):
    """inferencer.py:20-27: type ids follow set() iteration order."""
    type_to_index, index_to_type = {}, {}
    for word in set(vocab):
        index_to_type[len(index_to_type)] = word
        type_to_index[word] = len(type_to_index)
    return type_to_index, index_to_type


class Inferencer(object):
    def __init__(self, hyper_parameter_optimize_interval=10):
        # inferencer.py:32-39
        self._hyper_parameter_optimize_interval = hyper_parameter_optimize_interval
        assert self._hyper_parameter_optimize_interval > 0

    def _initialize(self, vocab, number_of_topics, alpha_alpha, alpha_beta):
        # inferencer.py:45-58
        self.parse_vocabulary(vocab)
        self._number_of_types = len(self._type_to_index)
        self._counter = 0
        self._number_of_topics = number_of_topics
        self._alpha_alpha = numpy.zeros(self._number_of_topics) + alpha_alpha
        self._alpha_beta = numpy.zeros(self._number_of_types) + alpha_beta

    def parse_vocabulary(self, vocab):
        # inferencer.py:60-67 (ids in set() order; run with PYTHONHASHSEED=0 for reproducible ids)
        self._type_to_index, self._index_to_type = parse_vocabulary(vocab)
        self._vocab = list(self._type_to_index.keys())    # a list, as under Python 2 (keeps the object picklable)

    def parse_data(self):
        raise NotImplementedError

    def learning(self):
        raise NotImplementedError

    def inference(self):
        raise NotImplementedError

    def export_beta(self, exp_beta_path, top_display=-1):
        raise NotImplementedError

    def export_gamma(self, exp_gamma_path, top_display=-1):
        raise NotImplementedError
