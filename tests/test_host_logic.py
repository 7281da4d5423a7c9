"""CPU tests of the host-side logic around the hot path: the py3 SCG replay, the flat
parameter transforms, and the world_size-2 (gloo) reduce plumbing of the N > 1 path."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from gparml_b200 import transforms as T
from gparml_b200.scg_adapted import SCG_adapted
from gparml_b200.synthetic import make_problem, split_rows
from oracle import gparml_oracle as O
from oracle_backend import OracleBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_transforms_match_oracle():
    x = np.array([-3.0, 0.0, 2.0])
    assert np.allclose([T.transform((0, None), v) for v in x], O.softplus(x))
    assert np.allclose([T.transform_grad((0, None), v) for v in x], O.softplus_grad(x))
    assert T.transform((None, None), -7.0) == -7.0 and T.transform_grad((None, None), 3.0) == 1
    assert np.allclose(T.transformVar_back(T.transformVar(x)), x)
    with pytest.raises(AssertionError):
        T.transformVar(np.array([37.0]))


def test_scg_minimises_quadratic_without_local_state():
    A = np.diag(np.arange(1.0, 7.0))
    b = np.arange(6.0)

    def f(x, iteration, step_size=0):
        return 0.5 * x @ A @ x - b @ x, A @ x - b
    x, flog, nev, status, _ = SCG_adapted(f, np.zeros(6), "unused", fixed_embeddings=True, maxiters=40, display=False,
                                          xtol=1e-12, ftol=1e-14, gtol=1e-14)
    assert np.allclose(x, np.linalg.solve(A, b), atol=1e-5)
    assert flog[-1] <= flog[0]


def test_scg_with_oracle_backend_decreases_bound_and_moves_embeddings():
    p = make_problem(60, 4, 2, 3, seed=21)
    shards = [dict(Y=p["Y"][lo:hi], X_mu=p["X_mu"][lo:hi], X_S=p["X_S"][lo:hi]) for lo, hi in split_rows(60, 2)]
    be = OracleBackend(shards, 4, 2)
    x0 = np.concatenate([p["Z"].ravel(), O.softplus_inv(np.array([1.0, 1.0, 1.0, 1.0]))])
    x, flog, nev, status, _ = SCG_adapted(be.f_and_gradf, x0, "unused", fixed_embeddings=False, maxiters=6,
                                          display=False, xtol=0, ftol=0, gtol=0, local_ops=be)
    assert flog[-1] < flog[0]
    assert not np.allclose(be.st[0]["X_mu"], shards[0]["X_mu"])
    assert status == "maxiter exceeded" and len(flog) == 7


def test_safe_wrapper_maps_failures_to_inf():
    from gparml_b200.scg_adapted import safe_f_and_grad_f

    def bad(x, it, st):
        raise np.linalg.LinAlgError("not PD")
    f, g = safe_f_and_grad_f(bad, np.zeros(3))
    assert f == np.inf and np.array_equal(g, np.ones(3))


GLOO_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np, torch
    from gparml_b200 import distributed as gd
    from gparml_b200.synthetic import make_problem
    from oracle import gparml_oracle as O
    rank, world, _ = gd.init_process_group("gloo")
    p = make_problem(41, 5, 2, 3, seed=4, generic_hypers=True)
    lo, hi = gd.shard_range(41, world, rank)
    S = O.softplus(p["X_S"])
    st = O.shard_statistics(p["Y"][lo:hi], p["X_mu"][lo:hi], S[lo:hi], p["Z"], p["sf2"], p["alpha"])
    packed = torch.from_numpy(np.concatenate([np.ravel(np.asarray(st[k], dtype=float)) for k in O.STAT_NAMES]))
    gd.allreduce_sum_(packed)
    full = O.shard_statistics(p["Y"], p["X_mu"], S, p["Z"], p["sf2"], p["alpha"])
    ref = np.concatenate([np.ravel(np.asarray(full[k], dtype=float)) for k in O.STAT_NAMES])
    err = float(np.max(np.abs(packed.numpy() - ref)) / np.max(np.abs(ref)))
    assert err < 1e-13, err

    class FakeCtx(object):          # one shard's partial inner products
        def scg_get_mu(self): return 1.5 + rank
        def scg_get_max_d(self, a): return a * (2.0 + rank)
    ops = gd.DistributedLocalOps(FakeCtx())
    assert ops.embeddings_get_grads_mu("x") == sum(1.5 + r for r in range(world))
    assert ops.embeddings_get_grads_max_d("x", 0.5) == 0.5 * (2.0 + world - 1)
    import torch.distributed as dist
    dist.barrier(); dist.destroy_process_group()
    print("rank %%d ok" %% rank)
""")


def test_world_size_2_gloo_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_bench_reference_arm_line_on_cpu():
    """`bench.py --impl reference` needs no GPU: it times the reference's own maps (staged oracle/_ref or
    /root/reference; the oracle port where neither exists) in a Pool on the host cores and prints ONE JSON line with
    the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c1",
                        "--steps", "1", "--warmup", "0", "--cpu-points", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ELBO+grad evals/sec" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    from oracle import ref_shim
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 exit without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c1", "--gpus", "2"],
                        capture_output=True, text=True, timeout=60, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


MR_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np
    from gparml_b200 import b200_MapReduce as mr
    work = %(work)r
    opts = dict(input=work + "/input", embeddings=work + "/emb", statistics=work + "/stat", tmp=work, M=3, Q=2, D=2,
                load=True, fixed_embeddings=False, init="PCA", b200_backend="gloo")
    opts = mr.init(opts)                       # counts the rows of this rank's files, sums over ranks; no device needed with load
    rank, world, tdev = mr.dist_info(opts)
    assert world == 2 and tdev is None
    assert opts["N"] == 5 + 7 + 4 and opts["b200_file_lengths"] == [5, 7, 4], (opts["N"], opts["b200_file_lengths"])
    s = mr._Session()
    mr._my_files(opts, s)                      # input file i belongs to rank i %% world
    assert s.file_index == list(range(rank, 3, 2)) and s.n_files_total == 3, s.file_index
    assert [os.path.basename(f) for f in s.files] == [["a.npy", "c.npy"], ["b_csv"]][rank]
    got = mr._bcast({"draw": np.random.RandomState(rank).rand(), "rank": rank}, opts)      # rank 0's host-side decisions everywhere
    assert got["rank"] == 0 and got["draw"] == np.random.RandomState(0).rand()
    import torch.distributed as dist
    dist.barrier(); dist.destroy_process_group()
    print("rank %%d ok" %% rank)
""")


def test_map_reduce_backend_shards_files_over_ranks_gloo(tmp_path):
    """Host logic of the one-process-per-GPU layout of b200_MapReduce on CPU (gloo, world_size 2): row counts summed
    over ranks, input file i -> rank i % world, CSV and .npy shards, rank 0's random decisions broadcast."""
    (tmp_path / "input").mkdir(); (tmp_path / "emb").mkdir(); (tmp_path / "stat").mkdir()
    rng = np.random.default_rng(0)
    np.save(str(tmp_path / "input" / "a.npy"), rng.standard_normal((5, 2)))
    np.savetxt(str(tmp_path / "input" / "b_csv"), rng.standard_normal((7, 2)), delimiter=",")
    np.save(str(tmp_path / "input" / "c.npy"), rng.standard_normal((4, 2)))
    script = tmp_path / "worker.py"
    script.write_text(MR_WORKER % {"root": ROOT, "work": str(tmp_path)})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29534", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_vectorised_transforms_match_scalar_ones():
    """The replayed driver transforms the flat parameter vector in one numpy pass; same values as the reference's
    per-element transform / transform_grad (supporting_functions.py:131-148)."""
    from gparml_b200 import parallel_GPLVM as drv
    bounds = [(None, None)] * 6 + [(0, None)] * 4
    opts = {"flat_positive": np.array([b == (0, None) for b in bounds])}
    x = np.random.default_rng(1).standard_normal(10) * 3
    assert np.array_equal(drv._transform_vec(opts, x), np.array([T.transform(b, v) for b, v in zip(bounds, x)]))
    assert np.array_equal(drv._transform_grad_vec(opts, x), np.array([T.transform_grad(b, v) for b, v in zip(bounds, x)], dtype=float))
    x[7] = 40.0
    with pytest.raises(AssertionError):
        drv._transform_vec(opts, x)
