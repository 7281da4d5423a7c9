import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def relerr(a, b):
    """max-norm relative error  max|a-b| / max|b|  (SURVEY.md 8d parity protocol)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = float(np.max(np.abs(b))) if b.size else 0.0
    if den == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b))) / den


def load_golden(name):
    """Load a fixture produced by oracle/gen_golden.py from the live reference."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    sizes = [int(s) for s in g["shard_sizes"]]
    shards, lo = [], 0
    for n in sizes:
        sh = dict(Y=g["Y"][lo:lo + n], X_mu=g["X_mu"][lo:lo + n], X_S=g["X_S"][lo:lo + n])
        if "d" in g:
            sh["d"] = g["d"][:, lo:lo + n]
        shards.append(sh)
        lo += n
    g["shards"] = shards
    g["sf2"] = float(g["sf2"]); g["beta"] = float(g["beta"]); g["step_size"] = float(g["step_size"])
    g["fixed_embeddings"] = bool(g["fixed_embeddings"])
    return g


GOLDEN_CASES = ["t5", "c1s", "c2s", "c3s", "c3u", "c4s"]
BIG_SAMPLE = (slice(None, None, 7), slice(None), slice(None, None, 11))


def check_against_golden(res, g, tol, what=""):
    """Compare an evaluation result dict (stats / global / grad_latest) with a fixture."""
    errs = {}
    for k in g:
        if k.startswith("stat_sample_"):
            name = k[len("stat_sample_"):]
            errs[name] = relerr(np.asarray(res["stats"][name])[BIG_SAMPLE], g[k])
        elif k.startswith("stat_"):
            name = k[len("stat_"):]
            errs[name] = relerr(res["stats"][name], g[k])
        elif k.startswith("glob_") and k != "glob_cond_Kmm":
            name = k[len("glob_"):]
            if name in res["global"]:
                errs[name] = relerr(res["global"][name], g[k])
        elif k.startswith("grad_latest_"):
            i = int(k[len("grad_latest_"):])
            errs[k] = relerr(res["grad_latest"][i], g[k])
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, "%s parity > %g: %r" % (what, tol, bad)
    return errs
