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

Scaling (the reference forks one mapper process per input file, local_MapReduce.py:134-137):

* one process, several GPUs (``options['b200_devices']``, default: all visible): shards are dealt
  round-robin over the devices; every map is *launched* on every shard before the host waits for any
  (``gparml_statistics_launch``), the reducer is one peer-memory kernel
  (``gparml_stats_allreduce_peers``) and the master step is replicated per shard context;
* one process per GPU (``python -m torch.distributed.run ... `` -- detected from ``WORLD_SIZE`` > 1,
  ``options['b200_distributed']`` = False disables it): input file ``i`` belongs to rank
  ``i % world``, the reducer becomes the in-process reduce followed by ONE all-reduce of the packed
  buffer (NCCL over NVLink), every rank runs the same master step and the same optimiser with
  bit-identical scalars; rank 0 alone writes the files in the statistics folder.
"""
import glob
import os
import random
import time
from os.path import basename

import numpy

from . import _lib
from . import engine
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
        self.files = []          # input files of THIS rank (all files in a single process)
        self.file_index = []     # their positions in the sorted list of all input files
        self.n_files_total = 0
        self.ctx = []
        self.root = None
        self.globals = None
        self.kept = None         # local indices of the shards that took part in the last statistics_MR
        self.rank, self.world, self.tdev = 0, 1, None


def dist_info(options=None):
    """(rank, world, torch device for small collectives or None).  world > 1 only under
    torch.distributed.run (one process per GPU) and unless options['b200_distributed'] is False."""
    if options is not None and not options.get("b200_distributed", True):
        return 0, 1, None
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return 0, 1, None
    import torch
    import torch.distributed as dist
    from . import distributed as gd
    if not dist.is_initialized():
        backend = (options or {}).get("b200_backend")
        if backend is None:
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
            backend = "nccl" if torch.cuda.device_count() >= local_world else "gloo"   # gloo: ranks share one GPU (tests)
        gd.init_process_group(backend)
    rank, world = dist.get_rank(), dist.get_world_size()
    tdev = None
    if dist.get_backend() == "nccl":
        tdev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    return rank, world, tdev


def _bcast(obj, options=None):
    """Rank 0's ``obj`` on every rank (host-side decisions that consume random numbers)."""
    rank, world, tdev = dist_info(options)
    if world == 1:
        return obj
    import torch.distributed as dist
    box = [obj if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, device=tdev)
    return box[0]


def _key(options):
    return os.path.abspath(options["embeddings"])


def _devices(options):
    n = _lib.load().gparml_device_count()
    if n <= 0:
        raise _lib.GparmlError("no CUDA device visible -- the b200 backend has no CPU path")
    rank, world, tdev = dist_info(options)
    if world > 1:                                     # one process per GPU (all ranks on device 0 under gloo)
        return [tdev.index if tdev is not None else 0]
    want = options.get("b200_devices")
    return list(want) if want else list(range(n))


def _load_csv(path):
    """One input shard: the reference's CSV (local_MapReduce.py:197), or -- an extension for large shards,
    where genfromtxt takes minutes -- the same matrix as a ``.npy`` file."""
    if path.endswith(".npy"):
        Y = numpy.load(path)
    else:
        Y = numpy.genfromtxt(path, delimiter=",")
    if Y.ndim == 1:                                   # local_MapReduce.py:198-199
        Y = numpy.atleast_2d(Y).T
    return numpy.ascontiguousarray(Y, dtype=numpy.float64)


def _count_rows(path):
    if path.endswith(".npy"):
        return int(numpy.load(path, mmap_mode="r").shape[0])
    n = 0
    with open(path) as f:
        for line in f:
            if line.strip():
                n += 1
    return n


def _my_files(options, s):
    """The input files this process maps: all of them, or every world-th one under torch.distributed.run."""
    s.rank, s.world, s.tdev = dist_info(options)
    every = sorted(glob.glob(options["input"] + "/*"))
    s.n_files_total = len(every)
    s.file_index = list(range(s.rank, len(every), s.world))
    s.files = [every[i] for i in s.file_index]
    if not s.files:
        raise ValueError("rank %d of %d has no input file: %d files in %s" % (s.rank, s.world, len(every), options["input"]))


def _new_context(options, s, i, **kw):
    devs = _devices(options)
    c = ShardContext(options["M"], options["Q"], options["D"], options["N"], device=devs[i % len(devs)],
                     fixed_beta=bool(options.get("fixed_beta")), **kw)
    if s.world > 1 or options.get("b200_stream") == "torch":
        c.use_torch_stream()                          # the all-reduce is issued by torch on its current stream
    return c


def _session(options):
    """Create (once) the device-resident shards: replaces the per-call loads of
    local_MapReduce.py:195-201 / 323-329."""
    k = _key(options)
    s = _sessions.get(k)
    if s is not None:
        return s
    s = _Session()
    _my_files(options, s)
    fixed = bool(options.get("fixed_embeddings"))
    for i, f in enumerate(s.files):
        Y = _load_csv(f)
        X_mu = load(options["embeddings"] + "/" + basename(f) + ".embedding.npy")
        X_S = load(options["embeddings"] + "/" + basename(f) + ".variance.npy")
        c = _new_context(options, s, i, fixed_embeddings=fixed)
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
    return session_of(folder).ctx


def session_of(folder):
    s = _sessions.get(os.path.abspath(folder))
    if s is None:
        raise ValueError("no device session for folder %r: run statistics_MR first" % folder)
    return s


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
    _my_files(options, s)
    for i, f in enumerate(s.files):
        Y = _load_csv(f)
        c = _new_context(options, s, i)
        c.upload_outputs(Y)
        s.ctx.append(c)
    s.root = s.ctx[0]
    seed = _bcast(int(numpy.random.randint(0, 2 ** 31 - 1)), options)   # follows numpy.random.seed like the reference's draws
    reduce_fn = None
    if s.world > 1:
        from . import distributed as gd
        reduce_fn = lambda a: gd.allreduce_numpy(a, device=s.tdev)      # noqa: E731
    # global row of each local shard's first point: the draws are functions of the global row
    offsets, lo = {}, 0
    for k, n in enumerate(options["b200_file_lengths"]):
        offsets[k] = lo
        lo += n
    row_offsets = [offsets[k] for k in s.file_index]
    if options["init"] == "PCA":
        init_device.pca(s.ctx, reduce_fn=reduce_fn)
    else:
        init_device.random_means(s.ctx, seed + 1, row_offsets)
    init_device.random_variances(s.ctx, seed, row_offsets)
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
    for c in s.ctx:                       # under torch.distributed.run: this rank's shards only (rank 0's Z is
        ctxs.append(c)                    # broadcast by init_statistics, no collective in here)
        n += c.n_local
        if n >= k:
            break
    cand = numpy.concatenate([c.download(_lib.A_X_MU, (c.n_local, c.Q)) for c in ctxs])
    return init_device.kmeans(ctxs, k, candidates=cand)[0]


def init(options):
    rank, world, tdev = dist_info(options)
    names = sorted(os.listdir(options["input"] + "/"))
    lengths = [0] * len(names)
    for i in range(rank, len(names), world):          # every rank counts its own files ...
        lengths[i] = _count_rows(options["input"] + "/" + names[i])
    if world > 1:                                     # ... and the counts are summed over ranks
        from . import distributed as gd
        lengths = [int(round(v)) for v in gd.allreduce_numpy(numpy.array(lengths, dtype=numpy.float64), device=tdev)]
    options["N"] = sum(lengths)
    options["b200_file_lengths"] = lengths
    if world > 1 and not options.get("b200_device_init", True) and not options["fixed_embeddings"] and not options["load"]:
        raise ValueError("under torch.distributed.run the embeddings are initialised on the devices (b200_device_init)")

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
    ``(list[(statistic, file)], mapper_times, reducer_times)`` like local_MapReduce.py:171.

    Nothing in here waits for a GPU: all maps are queued first (several GPUs then work concurrently),
    the reducer is queued behind them, and the first host wait is whoever reads a result."""
    global dropped_out_nodes, non_dropped_out_nodes
    s = _session(options)
    gs = _globals_for(options, s)
    idx = list(range(len(s.ctx)))
    scale = 1.0
    if options.get("drop_out_fraction", 0) > 0:                        # local_MapReduce.py:121-129
        n_all = s.n_files_total
        if s.rank == 0:
            drop = numpy.random.uniform(size=n_all) < options["drop_out_fraction"]
            dropped = [int(i) for i in numpy.arange(n_all)[drop]]
            kept = [int(i) for i in numpy.arange(n_all)[~drop]]
            if len(kept) == 0:
                # as in the reference the dropped list is left alone here, so the rescaling below is (1 + n) / 1
                kept = [random.randint(0, n_all - 1)]
        else:
            dropped = kept = None
        dropped, kept = _bcast((dropped, kept), options)
        dropped_out_nodes, non_dropped_out_nodes = dropped, kept
        idx = [k for k, g in enumerate(s.file_index) if g in kept]
        scale = float(len(kept) + len(dropped)) / len(kept)            # :263-264 (the reducer divides by kept/(kept+dropped))
    mapper_times = []
    for i in idx:
        t = time.time()
        _push(s.ctx[i], gs, options)
        s.ctx[i].statistics_launch()
        mapper_times.append(time.time() - t)
    t = time.time()
    if s.world == 1:
        engine.allreduce_contexts([s.ctx[i] for i in idx], scale)      # every kept context now holds the sums
        root = s.ctx[idx[0]]
    else:
        # this rank's shards first (a rank whose shards were all dropped contributes zeros), then ONE
        # all-reduce of the packed buffer over the ranks
        root = s.ctx[idx[0]] if idx else s.ctx[0]
        if idx:
            engine.allreduce_contexts([s.ctx[i] for i in idx], scale)
        else:
            _push(root, gs, options)
            root.set_stats_packed(numpy.zeros(root.stats_count))
        from . import distributed as gd
        gd.allreduce_sum_(root.stats_torch_view())
        for i in idx[1:]:
            s.ctx[i].stats_copy_from(root)
    s.root, s.kept = root, idx
    files = []
    if _write_files(options):
        if s.rank == 0:
            named = root.stats_named()
            for key in options["accumulated_statistics_names"]:
                name = options["statistics"] + "/accumulated_statistics_" + key + "_" + str(options["i"]) + ".npy"
                save(name, numpy.asarray(named[key]))
                files.append((key, name))
        else:
            files = [(key, None) for key in options["accumulated_statistics_names"]]
    else:
        files = [(key, None) for key in options["accumulated_statistics_names"]]
    return files, mapper_times, [time.time() - t]


# ------------------------------------------------------------------------------------------
# embeddings map (local_MapReduce.py:284-363); no reduce: gradients stay with their shard
# ------------------------------------------------------------------------------------------
def _launch_embeddings(options, s, gs):
    """Queue the replicated master step and the embeddings map on every shard of this process (all shards,
    dropped or not: local_MapReduce.py:292-294) without waiting for any of them."""
    times = []
    for i, c in enumerate(s.ctx):
        t = time.time()
        if s.kept is not None and i not in s.kept:
            _push(c, gs, options)        # a dropped-out shard has not seen this evaluation's globals yet
            c.stats_copy_from(s.root)
        c.global_step_begin()            # replicated master step: identical inputs, identical outputs
        c.embedding_grads()
        times.append(time.time() - t)
    return times


def embeddings_MR(options):
    s = _session(options)
    gs = _globals_for(options, s)
    times = _launch_embeddings(options, s, gs)
    for c in s.ctx:
        c.global_step_end()              # reports a failed pivot / variance out of range of this evaluation
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
    if _write_files(options) and s.rank == 0:
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
    stay on the devices.  Used by the py3 driver when ``options['b200_write_files']`` is False and
    always under torch.distributed.run (every rank must compute bit-identical scalars).

    The host queues the statistics maps of all shards, the reduce, the replicated master steps and the
    embeddings maps, and only then waits -- for F and the global gradient of the root context."""
    s = _session(options)
    s.globals = (options["i"], dict(global_statistics))
    statistics_MR(dict(options, b200_write_files=False))
    if options.get("fixed_embeddings"):
        return s.root.global_step()
    _launch_embeddings(options, s, s.globals[1])
    out = s.root.global_step_end()
    for c in s.ctx:
        if c is not s.root:
            c.global_step_end()
    return out


def write_evaluation_files(options, global_statistics):
    """The files of one evaluation, written from the device state after :func:`fast_evaluation`:
    ``global_statistics_*``, ``accumulated_statistics_*``, ``cache_Kmm[_inv]_*`` and ``partial_derivatives_*``
    (parallel_GPLVM.py:236-238,331; local_MapReduce.py:259,390-394) -- what ``predict.py`` and ``--load``
    read back from the ``'f'`` checkpoint.  Rank 0 only under torch.distributed.run."""
    s = _session(options)
    if s.rank != 0:
        return
    it, st, root = str(options["i"]), options["statistics"], s.root
    M, D = options["M"], options["D"]
    for key in global_statistics:
        save(st + "/global_statistics_" + key + "_" + it + ".npy", global_statistics[key])
    named = root.stats_named()
    for key in options["accumulated_statistics_names"]:
        save(st + "/accumulated_statistics_" + key + "_" + it + ".npy", numpy.asarray(named[key]))
    save(st + "/cache_Kmm_" + it + ".npy", root.download(_lib.A_KMM, (M, M)))
    save(st + "/cache_Kmm_inv_" + it + ".npy", root.download(_lib.A_KMM_INV, (M, M)))
    beta = float(numpy.squeeze(global_statistics["beta"]))
    pd = {"F": root.last_F, "dF_dsum_exp_K_ii": -0.5 * beta * D,                       # partial_terms.py:133-138
          "dF_dKmm": root.download(_lib.A_DF_DKMM, (M, M)),
          "dF_dsum_exp_K_miY": root.download(_lib.A_DF_DPSI1Y, (M, D)),
          "dF_dsum_exp_K_mi_K_im": root.download(_lib.A_DF_DPSI2, (M, M))}
    for key in options.get("partial_derivatives_names", pd):
        save(st + "/partial_derivatives_" + key + "_" + it + ".npy", pd[key])
