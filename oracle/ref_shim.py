"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference arithmetic.

Imports the reference's own ``kernels.py`` / ``kernel_exp.py`` /
``partial_terms.py`` from ``/root/reference`` (read-only, never copied into this
repo) under Python 3 so that the restated oracle (``oracle/gparml_oracle.py``)
and the golden vectors (``tests/golden``) can be pinned against the reference
itself.  ``/root/reference`` only exists inside the build container; parity tests (``-m gpu``) and
``smoke()`` never call :func:`load_reference` -- they use the committed fixtures.  The one consumer on the
GPU box is the CPU arm of ``bench.py``, which times the reference's own arithmetic from the staged copy
``oracle/_ref/`` (made by ``oracle/stage_ref.py``; git-ignored, so no reference source enters the history).

Two in-memory compatibility edits are needed; the files
on disk are untouched:

* ``builtins.xrange = range``  (partial_terms.py uses xrange throughout, e.g.
  partial_terms.py:200,224,250,261,278,294,383,416)
* ``kernels.py:20`` ``if ard==None`` and ``kernels.py:89`` ``if X2==None`` are
  ambiguous array truth tests under modern numpy -> rewritten to ``is None``
  in the source text before ``exec``.
"""
import builtins
import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# /root/reference in the build container; on the GPU box the git-ignored staging copy oracle/_ref/ that
# oracle/stage_ref.py made from it (three files, byte-identical, shipped with the snapshot, never committed)
REFERENCE_ROOT = os.environ.get("GPARML_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/partial_terms.py") else _STAGED)


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "partial_terms.py"))


_cache = {}


def load_reference():
    """Return ``(partial_terms_module, kernel_exp_module, kernels_module)``."""
    if "mods" in _cache:
        return _cache["mods"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not hasattr(builtins, "xrange"):
        builtins.xrange = range
    with open(os.path.join(REFERENCE_ROOT, "kernels.py")) as f:
        src = f.read()
    src = src.replace("if ard==None:", "if ard is None:")
    src = src.replace("if X2==None:", "if X2 is None:")
    kernels = types.ModuleType("kernels")
    kernels.__file__ = os.path.join(REFERENCE_ROOT, "kernels.py")
    exec(compile(src, kernels.__file__, "exec"), kernels.__dict__)
    saved = {k: sys.modules.get(k) for k in ("kernels", "kernel_exp", "partial_terms")}
    sys.modules["kernels"] = kernels
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        sys.modules.pop("kernel_exp", None)
        sys.modules.pop("partial_terms", None)
        kernel_exp = importlib.import_module("kernel_exp")
        partial_terms = importlib.import_module("partial_terms")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # do not leave the reference's top-level names shadowing anything else
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    # partial_terms holds direct references to its kernels/kernel_exp modules, so
    # popping them from sys.modules is safe.
    _cache["mods"] = (partial_terms, kernel_exp, kernels)
    return _cache["mods"]
