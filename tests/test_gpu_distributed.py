"""GPU test of the N > 1 path: one process per shard, packed statistics all-reduced with
torch.distributed.  On a one-GPU box both ranks share cuda:0 and the backend is gloo (NCCL
refuses two ranks on one device); with >= 2 GPUs the same worker runs over NCCL."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(backend, world, tmp_path, port):
    worker = os.path.join(ROOT, "tests", "dist_gpu_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), worker, backend, str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]


def _check(outs):
    from gparml_b200.synthetic import make_problem
    from oracle import c_oracle
    N, M, Q, D = 3001, 30, 4, 3
    p = make_problem(N, M, Q, D, seed=77, generic_hypers=True, with_direction=True)
    shard = dict(Y=p["Y"], X_mu=p["X_mu"], X_S=p["X_S"], d=p["d"])
    ref = c_oracle.evaluate([shard], p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=1e-3)
    flat_ref = np.concatenate([ref["global"]["grad_Z"].ravel(), [ref["global"]["grad_sf2"]], ref["global"]["grad_alpha"],
                               [ref["global"]["grad_beta"]]])
    for o in outs:
        assert float(o["F"]) == float(outs[0]["F"])                 # replicated master step: identical on every rank
        assert np.array_equal(o["flat"], outs[0]["flat"])
        assert abs(float(o["F"]) - ref["global"]["F"]) <= 1e-9 * abs(ref["global"]["F"])
        assert relerr(o["flat"], flat_ref) < 1e-9
        assert relerr(o["gl"], ref["grad_latest"][0][:, int(o["lo"]):int(o["hi"])]) < 1e-9
        assert float(o["kappa"]) == pytest.approx(float(np.sum(ref["grad_latest"][0] ** 2)), rel=1e-9)
    # device-side PCA / k-means with the partial sums all-reduced over ranks
    import scipy.cluster.vq as cl
    Y2 = make_problem(N, M, Q, 6, seed=78)["Y"]
    svd = np.linalg.svd(Y2 - Y2.mean(axis=0), full_matrices=False)
    ref_X = svd[0][:, :Q] / svd[0][:, :Q].std(axis=0)                # supporting_functions.py:116-120
    X0 = np.concatenate([o["X0"] for o in outs])
    assert relerr(X0 * np.sign(np.sum(X0 * ref_X, axis=0)), ref_X) < 1e-9
    ref_book, ref_dist = cl.kmeans(X0, p["Z"][:8] * 0.3)
    for o in outs:
        assert o["book"].shape == ref_book.shape and relerr(o["book"], ref_book) < 1e-9
        assert float(o["distortion"]) == pytest.approx(float(ref_dist), rel=1e-9)


def test_two_ranks_one_gpu_gloo(tmp_path):
    _check(_run("gloo", 2, tmp_path, 29541))


def test_ranks_over_nccl_if_multi_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _check(_run("nccl", min(n, 4), tmp_path, 29542))


# ----------------------------------------------------------------------------------------------------
# the drop-in API as the scalable path: parallel_GPLVM protocol + b200_MapReduce under torch.distributed.run
# ----------------------------------------------------------------------------------------------------
def _run_driver(backend, world, tmp_path, port, N=1000, M=2, Q=2, D=4, parts=4, iters=5, seed=1):
    from gparml_b200.synthetic import make_problem, split_rows
    p = make_problem(N, M, Q, D, seed=seed)
    for d in ("input", "embeddings", "statistics", "tmp"):
        (tmp_path / d).mkdir()
    for i, (lo, hi) in enumerate(split_rows(N, parts)):
        np.savetxt(str(tmp_path / "input" / ("easy_%d" % i)), p["Y"][lo:hi], delimiter=",", fmt="%.17g")
    worker = os.path.join(ROOT, "tests", "dist_driver_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), worker, backend, str(tmp_path),
                        str(M), str(Q), str(D), str(iters)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]


def _check_driver(outs, tmp_path, M, Q, iters):
    """Every rank followed the same optimiser trajectory, and it is the trajectory of the same SCG driver on the
    oracle backend started from the same initial state (local_MapReduce.py:115-171,284-308 semantics)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gparml_b200.scg_adapted import SCG_adapted
    from oracle import c_oracle
    from oracle_backend import OracleBackend
    for o in outs[1:]:
        assert np.array_equal(o["x"], outs[0]["x"]) and np.array_equal(o["flog"], outs[0]["flog"])
    assert sum(int(o["n_ctx"]) for o in outs) == len(os.listdir(str(tmp_path / "input")))
    names = sorted(os.listdir(str(tmp_path / "input")))
    init = str(tmp_path / "embeddings_init")
    shards = [dict(Y=np.genfromtxt(str(tmp_path / "input" / n), delimiter=","),
                   X_mu=np.load(os.path.join(init, n + ".embedding.npy")),
                   X_S=np.load(os.path.join(init, n + ".variance.npy"))) for n in names]
    be = OracleBackend(shards, M, Q, evaluate=c_oracle.evaluate)
    xo, flog_o, _, _, _ = SCG_adapted(be.f_and_gradf, outs[0]["x0"].copy(), "unused", False, display=False,
                                      maxiters=iters, xtol=0, ftol=0, gtol=0, local_ops=be)
    flog_g = outs[0]["flog"]
    assert len(flog_g) == iters + 1 and flog_g[-1] < flog_g[0]
    assert relerr(flog_g, np.array(flog_o)) < 1e-8
    assert relerr(outs[0]["x"], xo) < 1e-6
    for n, s in zip(names, be.st):            # every rank flushed the embeddings of its own shards
        assert relerr(np.load(str(tmp_path / "embeddings" / (n + ".embedding.npy"))), s["X_mu"]) < 1e-6
        assert relerr(np.load(str(tmp_path / "embeddings" / (n + ".grad_d.npy"))), s["d"]) < 1e-5
    # rank 0 wrote the 'f' checkpoint the prediction path / --load read
    for f in ("global_statistics_Z_f.npy", "accumulated_statistics_sum_exp_K_mi_K_im_f.npy", "cache_Kmm_inv_f.npy",
              "partial_derivatives_F_f.npy"):
        assert os.path.exists(str(tmp_path / "statistics" / f)), f
    assert float(np.load(str(tmp_path / "statistics" / "partial_derivatives_F_f.npy"))) == pytest.approx(-flog_g[-1], rel=1e-12)


def test_driver_two_ranks_one_gpu_gloo(tmp_path):
    """BASELINE config 1 (M=2 Q=2 D=4, N=1k in 4 shards, 5 SCG iterations) through parallel_GPLVM's protocol with
    two ranks (two shards each) sharing cuda:0."""
    _check_driver(_run_driver("gloo", 2, tmp_path, 29543), tmp_path, 2, 2, 5)


def test_driver_ranks_over_nccl_if_multi_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    _check_driver(_run_driver("nccl", world, tmp_path, 29544, N=4000, M=12, Q=3, D=5, parts=2 * world, seed=4),
                  tmp_path, 12, 3, 5)


def test_in_process_multi_gpu_shards_run_concurrently(tmp_path):
    """One host thread, shards dealt over all visible GPUs (or several shards on one): same results as one shard,
    and with >= 2 GPUs the maps overlap (the evaluation takes clearly less than the sum of its shards)."""
    import time
    import torch
    from gparml_b200.engine import ShardContext, evaluate
    from gparml_b200.synthetic import make_problem, split_rows
    ndev = torch.cuda.device_count()
    parts = max(2, min(ndev, 4))
    N, M, Q, D = 400000 * parts, 40, 6, 4        # ~4 ms of map per shard: well above launch latencies
    p = make_problem(N, M, Q, D, seed=31, generic_hypers=True)
    one = ShardContext(M, Q, D, N)
    one.upload_shard(p["Y"], p["X_mu"], p["X_S"])
    F1, g1 = evaluate([one], p["Z"], p["sf2"], p["alpha"], p["beta"])
    gl1 = one.grad_latest()
    ctxs = []
    for k, (lo, hi) in enumerate(split_rows(N, parts)):
        c = ShardContext(M, Q, D, N, device=k % ndev)
        c.upload_shard(p["Y"][lo:hi], p["X_mu"][lo:hi], p["X_S"][lo:hi])
        ctxs.append(c)
    evaluate(ctxs, p["Z"], p["sf2"], p["alpha"], p["beta"])
    for c in ctxs:
        c.synchronize()
    t0 = time.time()
    F, g = evaluate(ctxs, p["Z"], p["sf2"], p["alpha"], p["beta"])
    for c in ctxs:
        c.synchronize()
    t_all = time.time() - t0
    t0 = time.time()
    evaluate([one], p["Z"], p["sf2"], p["alpha"], p["beta"])
    one.synchronize()
    t_one = time.time() - t0
    assert abs(F - F1) <= 1e-11 * abs(F1) and relerr(g["flat"], g1["flat"]) < 1e-10
    assert relerr(np.concatenate([c.grad_latest() for c in ctxs], axis=1), gl1) < 1e-10
    st = [c.stats_packed() for c in ctxs]
    for s in st[1:]:
        assert np.array_equal(s, st[0])               # every context holds the reduced sums
    print("one shard %.2f ms, %d shards on %d GPU(s) %.2f ms" % (1e3 * t_one, parts, ndev, 1e3 * t_all))
    if ndev >= 2:
        assert t_all < 0.75 * t_one
    for c in ctxs + [one]:
        c.close()
