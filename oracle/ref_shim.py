"""TEST INFRASTRUCTURE ONLY -- runtime loader for the *unmodified* reference VB path.

Loads /root/reference/{inferencer,variational_bayes}.py (Python 2 sources) into
Python 3 modules at import time through a line-level source shim.  Nothing from
the reference is copied into this repository: the text is read where it lies,
rewritten in memory and exec'd.  The sources are searched in $PYLDA_REF, baseline/_ref/
(git-ignored staging copy made by stage(), the only way they reach the GPU box) and
/root/reference.  Callers: oracle/make_golden.py, the "not gpu" pin tests and the
`--impl reference` arm of bench.py.

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

_HERE = os.path.dirname(os.path.abspath(__file__))
# where the reference sources are looked for, in this order: $PYLDA_REF, baseline/_ref (git-ignored staging
# area that travels to the GPU box with the gpurun snapshot; filled by stage()), /root/reference
SEARCH = [p for p in (os.environ.get("PYLDA_REF"), os.path.join(_HERE, "..", "baseline", "_ref"), "/root/reference") if p]


def _find_root():
    for root in SEARCH:
        if os.path.isfile(os.path.join(root, "variational_bayes.py")) and os.path.isfile(os.path.join(root, "inferencer.py")):
            return os.path.abspath(root)
    return None


REFERENCE_ROOT = _find_root() or "/root/reference"


def available():
    return _find_root() is not None


def stage(dst=None):
    """Copy the two reference files of the path (inferencer.py, variational_bayes.py) from /root/reference into
    baseline/_ref/ -- git-ignored, never part of the repository's history, but shipped with the gpurun
    snapshot -- so that `bench.py --impl reference` can time the UNMODIFIED reference on the GPU box's host
    cores.  Returns the directory, or None when /root/reference is not there."""
    import shutil
    src = "/root/reference"
    if not os.path.isfile(os.path.join(src, "variational_bayes.py")):
        return None
    dst = dst or os.path.join(_HERE, "..", "baseline", "_ref")
    os.makedirs(dst, exist_ok=True)
    for name in ("inferencer.py", "variational_bayes.py"):
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
    return os.path.abspath(dst)


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
    root = _find_root()
    if root is None:
        raise RuntimeError("reference sources not found under any of %s" % SEARCH)
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
            path = os.path.join(root, name + ".py")
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
