"""Map-reduce backend interchangeable with the reference's ``local_MapReduce`` module
(``/root/reference/local_MapReduce.py``): same module-level functions, same ``options``
keys, same on-disk names, so ``parallel_GPLVM``'s ``map_reduce = <module>`` binding
(parallel_GPLVM.py:87-95) can point here (``options['parallel'] = 'b200'``).

What changes underneath: every input file becomes one device-resident shard context
(``engine.ShardContext``) created on first use -- the CSV is parsed once, not on every
evaluation (local_MapReduce.py:197,325) -- the mappers are CUDA kernels, the reducer is an
on-device sum of the packed statistics (NCCL all-reduce in the one-process-per-GPU layout of
``bench.py``), and the per-point gradients stay on the device for the optimiser's local-state
operations (``scg_adapted_b200_MapReduce``).  Files are still written where the reference's
callers read them back (``accumulated_statistics_*``, ``cache_*``, and -- on ``flush`` -- the
embeddings), unless ``options['b200_write_files']`` is False.
"""
import glob
import os
import time
from os.path import basename

import numpy

from . import _lib
from . import init_device
from . import partial_terms as pt
from . import transforms as sp
from .engine import ShardContext

# one session per embeddings folder (the reference keeps the same state in files there)
_sessions = {}
# Keep the dropped out nodes between the two MRs like local_MapReduce.py:108-109
dropped_out_nodes = []
non_dropped_out_nodes = []


class _Session(object):
    def __init__(self):
        self.files = []
        self.ctx = []
        self.root = None
        self.globals = None
        self.kept = None


def _key(options):
    return os.path.abspath(options["embeddings"])


def _devices(options):
    n = _lib.load().gparml_device_count()
    if n <= 0:
        raise _lib.GparmlError("no CUDA device visible -- the b200 backend has no CPU path")
    want = options.get("b200_devices")
    return list(want) if want else list(range(n))


def _load_csv(path):
    Y = numpy.genfromtxt(path, delimiter=",")
    if Y.ndim == 1:                                   # local_MapReduce.py:198-199
        Y = numpy.atleast_2d(Y).T
    return numpy.ascontiguousarray(Y)


def _session(options):
    """Create (once) the device-resident shards: replaces the per-call loads of
    local_MapReduce.py:195-201 / 323-329."""
    k = _key(options)
    s = _sessions.get(k)
    if s is not None:
        return s
    s = _Session()
    s.files = sorted(glob.glob(options["input"] + "/*"))
    devs = _devices(options)
    fixed = bool(options.get("fixed_embeddings"))
    for i, f in enumerate(s.files):
        Y = _load_csv(f)
        X_mu = load(options["embeddings"] + "/" + basename(f) + ".embedding.npy")
        X_S = load(options["embeddings"] + "/" + basename(f) + ".variance.npy")
        c = ShardContext(options["M"], options["Q"], options["D"], options["N"], device=devs[i % len(devs)],
                         fixed_embeddings=fixed, fixed_beta=bool(options.get("fixed_beta")))
        c.upload_shard(Y, X_mu, X_S)
        d_name = options["embeddings"] + "/" + basename(f) + ".grad_d.npy"
        if not fixed and exists(d_name):
            c.upload(_lib.A_GRAD_D, load(d_name))
        s.ctx.append(c)
    s.root = s.ctx[0]
    _sessions[k] = s
    return s


def close(options=None):
    """Release the device contexts (all sessions, or the one of ``options``)."""
    keys = list(_sessions) if options is None else [_key(options)]
    for k in keys:
        s = _sessions.pop(k, None)
        if s:
            for c in s.ctx:
                c.close()


def session_contexts(folder):
    """The shard contexts whose local state lives 'in' ``folder`` (used by the optimiser's
    local-state module, which the reference addresses by folder name too)."""
    s = _sessions.get(os.path.abspath(folder))
    if s is None:
        raise ValueError("no device session for folder %r: run statistics_MR first" % folder)
    return s.ctx


# ------------------------------------------------------------------------------------------
# init (local_MapReduce.py:27-104): same files; PCA / random draws on the device
# ------------------------------------------------------------------------------------------
def _device_init(options):
    """Replaces the master-side 'load ALL data' PCA of local_MapReduce.py:52-65 and the draws of
    :86-93: every CSV is parsed once into its shard context (which statistics_MR then reuses),
    the PCA runs on the shards' partial sums, and the files the reference writes
    (``.embedding.npy``, ``.variance.npy``) are flushed from the device."""
    close(options)
    s = _Session()
    s.files = sorted(glob.glob(options["input"] + "/*"))
    devs = _devices(options)
    for i, f in enumerate(s.files):
        Y = _load_csv(f)
        c = ShardContext(options["M"], options["Q"], options["D"], options["N"], device=devs[i % len(devs)],
                         fixed_beta=bool(options.get("fixed_beta")))
        c.upload_outputs(Y)
        s.ctx.append(c)
    s.root = s.ctx[0]
    seed = int(numpy.random.randint(0, 2 ** 31 - 1))     # follows numpy.random.seed like the reference's draws
    if options["init"] == "PCA":
        init_device.pca(s.ctx)
    else:
        init_device.random_means(s.ctx, seed + 1)
    init_device.random_variances(s.ctx, seed)
    _sessions[_key(options)] = s
    for f, c in zip(s.files, s.ctx):
        base = options["embeddings"] + "/" + basename(f)
        for ext in (".embedding.npy", ".variance.npy", ".grad_d.npy"):
            remove(base + ext)
        save(base + ".embedding.npy", c.download(_lib.A_X_MU, (c.n_local, c.Q)))
        save(base + ".variance.npy", c.download(_lib.A_X_S, (c.n_local, c.Q)))


def kmeans(options, k):
    """Device k-means over the embeddings the reference clusters (the first shards holding at
    least ``k`` points, parallel_GPLVM.py:170-181); returns the code book like
    ``scipy.cluster.vq.kmeans(...)[0]``."""
    s = _session(options)
    ctxs, n = [], 0
    for c in s.ctx:
        ctxs.append(c)
        n += c.n_local
        if n >= k:
            break
    cand = numpy.concatenate([c.download(_lib.A_X_MU, (c.n_local, c.Q)) for c in ctxs])
    return init_device.kmeans(ctxs, k, candidates=cand)[0]


def init(options):
    names = os.listdir(options["input"] + "/")
    lengths = []
    for name in names:
        n = 0
        with open(options["input"] + "/" + name) as f:
            for line in f:
                if line.strip():
                    n += 1
        lengths.append(n)
    options["N"] = sum(lengths)

    if not options["fixed_embeddings"] and not options["load"] and options.get("b200_device_init", True):
        if options["init"] not in ("PCA", "random"):
            raise ValueError("init=%r is not available in the b200 backend (PCA or random)" % options["init"])
        _device_init(options)
        return options
    if not options["fixed_embeddings"] and not options["load"]:
        X = None
        if options["init"] == "PCA":
            Y = numpy.concatenate([_load_csv(options["input"] + "/" + name) for name in names])
            X = sp.PCA(Y, options["Q"])
        elif options["init"] != "random":
            raise ValueError("init=%r is not available in the b200 backend (PCA or random)" % options["init"])
        lo = 0
        for name, n in zip(names, lengths):
            e_name = options["embeddings"] + "/" + name + ".embedding.npy"
            v_name = options["embeddings"] + "/" + name + ".variance.npy"
            remove(e_name)
            save(e_name, X[lo:lo + n, :] if X is not None else numpy.random.randn(n, options["Q"]))
            remove(v_name)
            save(v_name, sp.transformVar_back(numpy.clip(numpy.ones((n, options["Q"])) * 0.5
                                                         + 0.01 * numpy.random.randn(n, options["Q"]), 0.001, 1)))
            lo += n
    if options["fixed_embeddings"]:
        for name, n in zip(names, lengths):
            e_name = options["embeddings"] + "/" + name + ".embedding.npy"
            if not exists(e_name):
                raise Exception("No embedding file " + e_name)
            save(options["embeddings"] + "/" + name + ".variance.npy", numpy.zeros((n, options["Q"])))
    close(options)          # a fresh init invalidates any device copy of the old files
    return options


# ------------------------------------------------------------------------------------------
# statistics map-reduce (local_MapReduce.py:115-277)
# ------------------------------------------------------------------------------------------
def _write_files(options):
    return options.get("b200_write_files", True)


def _globals_for(options, s):
    if s.globals is not None and s.globals[0] == options["i"]:
        return s.globals[1]
    gs = {}
    for key in options["global_statistics_names"]:
        gs[key] = load(options["statistics"] + "/global_statistics_" + key + "_" + str(options["i"]) + ".npy")
    return gs


def _push(c, gs, options):
    c.set_globals(gs["Z"], float(numpy.squeeze(gs["sf2"])), numpy.squeeze(gs["alpha"]), float(numpy.squeeze(gs["beta"])))
    c.set_step(0.0 if options.get("fixed_embeddings") else float(options.get("step_size", 0) or 0))


def statistics_MR(options):
    """Runs the statistics map on every (kept) shard, reduces on the device and returns
    ``(list[(statistic, file)], mapper_times, reducer_times)`` like local_MapReduce.py:171."""
    global dropped_out_nodes, non_dropped_out_nodes
    s = _session(options)
    gs = _globals_for(options, s)
    idx = list(range(len(s.ctx)))
    scale = 1.0
    if options.get("drop_out_fraction", 0) > 0:                        # local_MapReduce.py:121-129
        drop = numpy.random.uniform(size=len(idx)) < options["drop_out_fraction"]
        dropped_out_nodes = [i for i in idx if drop[i]]
        non_dropped_out_nodes = [i for i in idx if not drop[i]]
        if not non_dropped_out_nodes:
            non_dropped_out_nodes = [int(numpy.random.randint(0, len(idx)))]
            dropped_out_nodes = [i for i in idx if i not in non_dropped_out_nodes]
        idx = list(non_dropped_out_nodes)
        scale = float(len(idx) + len(dropped_out_nodes)) / len(idx)    # :263-264
    mapper_times = []
    for i in idx:
        t = time.time()
        _push(s.ctx[i], gs, options)
        s.ctx[i].statistics()
        mapper_times.append(time.time() - t)
    t = time.time()
    root = s.ctx[idx[0]]
    for n, i in enumerate(idx[1:]):
        root.stats_add_any(s.ctx[i], scale if n == len(idx) - 2 else 1.0)
    if len(idx) == 1 and scale != 1.0:
        root.stats_add_any(root, 0.5 * scale)                          # (x + x) * scale / 2
    s.root, s.kept = root, idx
    files = []
    if _write_files(options):
        named = root.stats_named()
        for key in options["accumulated_statistics_names"]:
            name = options["statistics"] + "/accumulated_statistics_" + key + "_" + str(options["i"]) + ".npy"
            save(name, numpy.asarray(named[key]))
            files.append((key, name))
    else:
        files = [(key, None) for key in options["accumulated_statistics_names"]]
    return files, mapper_times, [time.time() - t]


# ------------------------------------------------------------------------------------------
# embeddings map (local_MapReduce.py:284-363); no reduce: gradients stay with their shard
# ------------------------------------------------------------------------------------------
def embeddings_MR(options):
    s = _session(options)
    gs = _globals_for(options, s)
    times = []
    for i, c in enumerate(s.ctx):        # all shards, dropped or not (local_MapReduce.py:292-294)
        t = time.time()
        if s.kept is not None and i not in s.kept:
            _push(c, gs, options)        # a dropped-out shard has not seen this evaluation's globals yet
        if c is not s.root:
            c.stats_copy_from(s.root)
        c.global_step()                  # replicated master step: identical inputs, identical outputs
        c.embedding_grads()
        times.append(time.time() - t)
    if _write_files(options) and options.get("b200_write_grad_files", False):
        for f, c in zip(s.files, s.ctx):
            save(options["embeddings"] + "/" + basename(f) + ".grad_latest.npy", c.grad_latest())
    return times


def flush(options):
    """Write the device-resident local state back to the reference's files
    (``.embedding.npy``, ``.variance.npy`` in the unconstrained domain, ``.grad_*.npy``):
    the checkpoint predict.py / --load / tools/show_embeddings.py read."""
    s = _session(options)
    for f, c in zip(s.files, s.ctx):
        base = options["embeddings"] + "/" + basename(f)
        n, Q = c.n_local, c.Q
        save(base + ".embedding.npy", c.download(_lib.A_X_MU, (n, Q)))
        save(base + ".variance.npy", c.download(_lib.A_X_S, (n, Q)))
        if not options.get("fixed_embeddings"):
            for nm, aid in (("latest", _lib.A_GRAD_LATEST), ("new", _lib.A_GRAD_NEW), ("old", _lib.A_GRAD_OLD),
                            ("d", _lib.A_GRAD_D)):
                save(base + ".grad_" + nm + ".npy", c.download(aid, (2, n, Q)))


# ------------------------------------------------------------------------------------------
# supporting functions (local_MapReduce.py:370-409)
# ------------------------------------------------------------------------------------------
def save(file_name, obj):
    numpy.save(file_name, obj)


def load(file_name):
    return numpy.load(file_name)


def exists(file_name):
    return os.path.exists(file_name)


def remove(file_name):
    if exists(file_name):
        os.remove(file_name)


def cache(options, global_statistics):
    """Kmm and Kmm^-1 once per evaluation (local_MapReduce.py:383-394); also remembers the
    globals of iteration ``options['i']`` so that the mappers need not re-read them."""
    s = _session(options)
    s.globals = (options["i"], dict(global_statistics))
    root = s.ctx[0]
    _push(root, global_statistics, options)
    root.update_global_statistics()
    if _write_files(options):
        M = options["M"]
        save(options["statistics"] + "/cache_Kmm_" + str(options["i"]) + ".npy", root.download(_lib.A_KMM, (M, M)))
        save(options["statistics"] + "/cache_Kmm_inv_" + str(options["i"]) + ".npy", root.download(_lib.A_KMM_INV, (M, M)))


def load_cache(options, partial_terms):
    Kmm = load(options["statistics"] + "/cache_Kmm_" + str(options["i"]) + ".npy")
    Kmm_inv = load(options["statistics"] + "/cache_Kmm_inv_" + str(options["i"]) + ".npy")
    partial_terms.set_global_statistics(Kmm, Kmm_inv)


def load_partial_terms(options, global_statistics):
    return pt.partial_terms(global_statistics["Z"], float(numpy.squeeze(global_statistics["sf2"])),
                            numpy.squeeze(global_statistics["alpha"]), float(numpy.squeeze(global_statistics["beta"])),
                            options["M"], options["Q"], options["N"], options["D"], update_global_statistics=False,
                            device=_devices(options)[0])


def fast_evaluation(options, global_statistics):
    """The whole of SURVEY.md 3.2 steps 5-9 without any file transport: returns
    ``(F, grad dict)`` with the global gradient in the positive domain; per-point gradients
    stay on the devices.  Used by the py3 driver when ``options['b200_write_files']`` is False."""
    s = _session(options)
    s.globals = (options["i"], dict(global_statistics))
    statistics_MR(dict(options, b200_write_files=False))
    F, grad = s.root.global_step()
    if not options.get("fixed_embeddings"):
        embeddings_MR(dict(options, b200_write_files=False))
    return F, grad
