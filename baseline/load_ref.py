"""Loader for the UNMODIFIED reference files under baseline/_ref/ (CPU arm of bench.py; never imported by the product).

`__graft_entry__.build()` copies Scribe/{information_estimators,causal_network,Scribe}.py from /root/reference into
baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).  They are imported here as SURVEY.md Appendix B describes:
empty stub modules for matplotlib (information_estimators.py:9 imports pyplot at module scope; only `sc` uses it), a synthetic
package `Scribe` whose __path__ is baseline/_ref (skips Scribe/__init__.py, which needs seaborn/statsmodels).  Nothing is edited.
"""
import importlib
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("information_estimators.py", "causal_network.py", "Scribe.py")


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def load():
    """(information_estimators, Scribe) modules of the reference, or None when baseline/_ref is absent."""
    if not available():
        return None
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "Scribe" not in sys.modules or getattr(sys.modules["Scribe"], "__path__", None) != [REF_DIR]:
        pkg = types.ModuleType("Scribe"); pkg.__path__ = [REF_DIR]; sys.modules["Scribe"] = pkg
        for sub in ("Scribe.information_estimators", "Scribe.Scribe"):
            sys.modules.pop(sub, None)
    ie = importlib.import_module("Scribe.information_estimators")
    sc = importlib.import_module("Scribe.Scribe")
    return ie, sc
