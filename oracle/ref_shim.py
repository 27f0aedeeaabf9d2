"""TEST INFRASTRUCTURE ONLY -- runtime loader for the *unmodified* reference VB path.

Loads /root/reference/{inferencer,variational_bayes}.py (Python 2 sources) into
Python 3 modules at import time through a line-level source shim.  Nothing from
the reference is copied into this repository: the text is read where it lies,
rewritten in memory and exec'd.  This only works where /root/reference exists
(the build container); the GPU box has no reference, so only
oracle/make_golden.py and the "not gpu" pin tests may call this.

Shim steps:
  1. stub `nltk` (imported at variational_bayes.py:10 / inferencer.py:8, unused on the path)
  2. scipy.misc.logsumexp := scipy.special.logsumexp (used variational_bayes.py:155,182,332)
  3. `print X` -> `print(X)`, xrange -> range, dict.keys()/values() -> list(...), cPickle -> pickle
  4. exec into modules named `inferencer` / `variational_bayes`
"""
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("PYLDA_REF", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "variational_bayes.py"))


_PRINT = re.compile(r"^(\s*)print\s+(?!\()(.*?);?\s*$")


def _to_py3(text):
    out = []
    for line in text.split("\n"):
        m = _PRINT.match(line)
        if m and not line.lstrip().startswith("#"):
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = line.replace("xrange(", "range(")
        line = line.replace("numpy.array(document_word_dict.keys())",
                            "numpy.array(list(document_word_dict.keys()))")
        line = line.replace("numpy.array(document_word_dict.values())",
                            "numpy.array(list(document_word_dict.values()))")
        line = line.replace("cPickle", "pickle")
        out.append(line)
    return "\n".join(out)


def load():
    """Return (inferencer_module, variational_bayes_module) built from the reference sources."""
    if not available():
        raise RuntimeError("reference sources not found under %s" % REFERENCE_ROOT)
    import scipy
    import scipy.special
    if "nltk" not in sys.modules:
        sys.modules["nltk"] = types.ModuleType("nltk")
    if "scipy.misc" not in sys.modules:
        misc = types.ModuleType("scipy.misc")
        sys.modules["scipy.misc"] = misc
        scipy.misc = misc
    import scipy.misc
    scipy.misc.logsumexp = scipy.special.logsumexp

    mods = {}
    saved = {n: sys.modules.get(n) for n in ("inferencer", "variational_bayes")}
    try:
        for name in ("inferencer", "variational_bayes"):
            path = os.path.join(REFERENCE_ROOT, name + ".py")
            with open(path, "r") as f:
                src = _to_py3(f.read())
            mod = types.ModuleType(name)
            mod.__file__ = path
            mod.__name__ = "_ref_" + name  # keeps the `if __name__ == "__main__"` stubs inert
            sys.modules[name] = mod        # so `from inferencer import ...` inside the reference resolves
            exec(compile(src, path, "exec"), mod.__dict__)
            mods[name] = mod
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return mods["inferencer"], mods["variational_bayes"]
