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
