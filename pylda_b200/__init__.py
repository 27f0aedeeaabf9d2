"""B200-native variational-Bayes E-step for LDA behind PyLDA's class API.

Host side mirrors /root/reference/{inferencer,variational_bayes,launch_train,launch_test}.py;
the E-step itself runs in pylda_b200/csrc (hand-written sm_100a CUDA behind a C ABI,
include/pylda_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
