#!/usr/bin/env python
"""Benchmark of the map-reduce variational-bound hot path (BASELINE.json metric:
ELBO+gradient evaluations per second at N=1M, M=100, Q=10, D=10 on 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU arithmetic on the host cores

One "step" = one full evaluation: globals host->device, prep_points,
psi2_stats, psi1_stats, [NCCL all-reduce of the packed sums], global_step (F and the global
gradient back on the host), embed_grads.  N > 1 shards the points over the ranks
(strong scaling: the problem size is fixed by the metric).

``value``          : device-resident shard (Y, X_mu, X_S uploaded once before the timed region), C-ABI calls.
``e2e``            : the same evaluation through the host-buffer API every step: the shard is copied
                     from pinned host memory (what ``partial_terms.set_data`` receives in the
                     reference's mappers, local_MapReduce.py:197-224) and the per-point gradients
                     are copied back (the ``.grad_latest.npy`` the reference writes, :359-360).
``through_driver`` : the same evaluation through the reference's own interface -- the replayed
                     ``parallel_GPLVM.likelihood_and_gradient(x, i, step)`` callback on the ``b200_MapReduce``
                     backend (device-resident shards, ``b200_write_files=False``).
``other_configs``  : BASELINE configs 5, 4 and 2 measured in the same run (fewer steps), with their own clocks.

Synthetic data is generated in 8 independent row blocks (gparml_b200/synthetic.py) so that every rank builds
only its rows and the problem -- hence F -- is the same for 1, 2, 4 and 8 ranks; F is checked against the
committed value in tests/golden/bench_expected_F.json.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# stdout carries exactly one JSON line (rank 0): NCCL's own banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gparml_b200.synthetic import CONFIGS, ROW_BLOCKS, block_problem_globals, block_problem_rows  # noqa: E402

E_EXP = 18  # FP64-pipe instructions of exp()
METRIC = "ELBO+grad evals/sec"
UNIT = "evals/s"
EXPECTED_F = os.path.join(ROOT, "tests", "golden", "bench_expected_F.json")


def algorithmic_ops(N, M, Q, D, fixed):
    """FP64-pipe lane-ops of one evaluation over N points."""
    P = M * (M + 1) // 2
    if fixed:
        return N * (P * (6 * Q + E_EXP + 2) + M * (6 * Q + E_EXP + 1 + D + 2 * Q * D))
    return N * (P * (12 * Q + 2 * E_EXP + 5) + M * (12 * Q + 2 * E_EXP + 4 + 3 * D + 2 * Q * D))


# ----------------------------------------------------------------------------------------
# CPU arm: the reference's own arithmetic (unmodified partial_terms.py / kernel_exp.py / kernels.py through
# oracle/ref_shim.py, mapper bodies replayed by oracle/ref_harness.py) hosted in a multiprocessing.Pool with
# one shard per worker, as local_MapReduce.py:134 does.  Where the reference files are not available
# (neither /root/reference nor the staged oracle/_ref), the numpy port oracle/gparml_oracle.py runs instead.
# ----------------------------------------------------------------------------------------
def _cpu_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"


def _cpu_stats_worker(a):
    kind, sh, Z, sf2, alpha, beta, N, D, fixed = a
    t = time.time()
    if kind == "reference":
        from oracle import ref_harness as H
        st = H.map_statistics(sh, Z, sf2, alpha, beta, N, D, 0.0, fixed)
    else:
        from oracle import gparml_oracle as O
        mu, S, _ = O.effective_embedding(sh["X_mu"], sh["X_S"], None, 0.0, fixed)
        st = O.shard_statistics_chunked(sh["Y"], mu, S, Z, sf2, alpha, chunk=256)
    return st, time.time() - t


def _cpu_embed_worker(a):
    kind, sh, stats, Z, sf2, alpha, beta, N, D, G1, G2 = a
    t = time.time()
    if kind == "reference":
        from oracle import ref_harness as H
        g = H.map_embeddings(sh, stats, Z, sf2, alpha, beta, N, D, 0.0)
    else:
        from oracle import gparml_oracle as O
        mu, S, sraw = O.effective_embedding(sh["X_mu"], sh["X_S"], None, 0.0, False)
        gm, gs = O.embedding_grads(sh["Y"], mu, S, Z, sf2, alpha, G1, G2)
        g = -np.array([gm, gs * O.softplus_grad(sraw)])
    return float(np.abs(g).sum()), time.time() - t


class CpuArm(object):
    def __init__(self, cfg_name, cores, pts_per_worker):
        import multiprocessing
        from oracle import ref_shim  # noqa: F401  (the CPU arm is the one place bench.py may use oracle/)
        self.kind = "reference" if ref_shim.reference_available() else "port"
        self.g = block_problem_globals(cfg_name)
        k = CONFIGS[cfg_name]
        self.k, self.cores, self.pts = k, cores, pts_per_worker
        n = cores * pts_per_worker
        bs = k["N"] // ROW_BLOCKS
        rows = block_problem_rows(cfg_name, 0, bs * ((n + bs - 1) // bs), with_direction=False)
        self.n = n
        self.shards = [dict(Y=rows["Y"][i * pts_per_worker:(i + 1) * pts_per_worker],
                            X_mu=rows["X_mu"][i * pts_per_worker:(i + 1) * pts_per_worker],
                            X_S=rows["X_S"][i * pts_per_worker:(i + 1) * pts_per_worker]) for i in range(cores)]
        # spawn (not fork): the parent may already hold a CUDA context; thread env is inherited
        saved = {v: os.environ.get(v) for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
        _cpu_init()
        self.pool = multiprocessing.get_context("spawn").Pool(cores, initializer=_cpu_init)
        for v, old in saved.items():
            if old is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = old

    def step(self):
        """One evaluation of the sample; returns (wall seconds of the two maps, wall seconds of the master step)."""
        g, k = self.g, self.k
        Z, sf2, alpha, beta, D, fixed = g["Z"], g["sf2"], g["alpha"], g["beta"], k["D"], k["fixed_embeddings"]
        t0 = time.time()
        res = self.pool.map(_cpu_stats_worker, [(self.kind, s, Z, sf2, alpha, beta, self.n, D, fixed) for s in self.shards])
        t_map = time.time() - t0
        t1 = time.time()
        if self.kind == "reference":
            from oracle import ref_harness as H
            stats = H.reduce_statistics([r[0] for r in res])
            gl = H.master_step(stats, Z, sf2, alpha, beta, self.n, D)
        else:
            from oracle import gparml_oracle as O
            stats = O.reduce_statistics([r[0] for r in res])
            gl = O.global_step(stats, Z, sf2, alpha, beta, self.n)
        t_glob = time.time() - t1
        if not fixed:
            t2 = time.time()
            self.pool.map(_cpu_embed_worker, [(self.kind, s, stats, Z, sf2, alpha, beta, self.n, D, gl["dF_dsum_exp_K_miY"],
                                               gl["dF_dsum_exp_K_mi_K_im"]) for s in self.shards])
            t_map += time.time() - t2
        return t_map, t_glob

    def close(self):
        self.pool.close()
        self.pool.join()

    def evals_per_s(self, t_map, t_glob):
        """Linear extrapolation of the maps to the full N (exactly linear: one Python iteration
        per point, partial_terms.py:46,200,278,383,416) plus the master step once."""
        full = t_map * (self.k["N"] / float(self.n)) + t_glob
        return 1.0 / full

    def describe(self):
        what = ("the reference's own partial_terms.py / kernel_exp.py / kernels.py (unmodified, oracle/ref_shim.py), mapper "
                "bodies of local_MapReduce.py:183-248,310-363 replayed on in-memory arrays" if self.kind == "reference"
                else "numpy port of the reference maps (oracle/gparml_oracle.py)")
        return ("%s in a %d-process Pool, %d points per worker (%d of %d points), one evaluation per step, map time "
                "extrapolated linearly to N; CSV parsing and .npy transport excluded"
                % (what, self.cores, self.pts, self.n, self.k["N"]))


def default_cpu_points(cfg_name, leg):
    M = CONFIGS[cfg_name]["M"]
    if M >= 400:
        return 8
    return 256


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    pts = args.cpu_points or default_cpu_points(args.config, "arm")
    cb = CpuArm(args.config, cores, pts)
    for _ in range(args.warmup):
        cb.step()
    t0 = time.time()
    tm = tg = 0.0
    for _ in range(args.steps):
        a, b = cb.step()
        tm += a
        tg += b
    wall = time.time() - t0
    cb.close()
    val = cb.evals_per_s(tm / args.steps, tg / args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.config, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": cb.kind, "sample": cb.describe()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample_wall_s_per_step": wall / args.steps,
    }
    print(json.dumps(line))
    return 0


def workload_config(cfg_name, gpus):
    k = CONFIGS[cfg_name]
    return {"workload": "%s: %s N=%d M=%d Q=%d D=%d, one ELBO+gradient evaluation" % (
                cfg_name, "sparse GP regression (fixed embeddings)" if k["fixed_embeddings"] else "Bayesian GPLVM",
                k["N"], k["M"], k["Q"], k["D"]),
            "N": k["N"], "M": k["M"], "Q": k["Q"], "D": k["D"], "shards": gpus,
            "parallelism": "points sharded over %d GPU(s); one packed fp64 all-reduce per evaluation" % gpus,
            "l2": "per-step inputs exceed L2 (point records %.0f MB per GPU)" % (
                k["N"] / gpus * (((3 * k["Q"] + 2) & ~1) * 8 * 2) / 1e6)}


# ----------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md)
# ----------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("uuid,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid = uuid
        self.lines = []          # (host receive time, csv line)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def window(self, t0, t1):
        """Summary of the samples received in the wall-clock window [t0, t1] (a timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)          # let the sample that covers the end of the window arrive
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [(t, ln) for t, ln in list(self.lines) if t >= t0 and t <= t1 + 0.15]
        for _, ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8 or (self.uuid and self.uuid not in f[0]):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()


# ----------------------------------------------------------------------------------------
# one configuration on this rank's GPU
# ----------------------------------------------------------------------------------------
class Env(object):
    """torch / torch.distributed handles shared by the configurations of one run."""
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- gparml_b200 has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # NCCL prints its version banner on stdout when the communicator is created; stdout must carry
            # exactly one JSON line, so file descriptor 1 points at stderr until the first collective is done
            sys.stdout.flush()
            saved_fd = os.dup(1)
            try:
                os.dup2(2, 1)
                dist.init_process_group("nccl", device_id=self.dev)
                warm = torch.zeros(1, device=self.dev)
                dist.all_reduce(warm)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_fd, 1)
                os.close(saved_fd)
        uuid = ""
        try:
            uuid = str(torch.cuda.get_device_properties(self.local_rank).uuid)
        except Exception:
            pass
        self.sampler = ClockSampler(uuid) if self.rank == 0 else None
        if self.sampler:
            self.sampler.start()

    def timed(self, fn, steps, per_step=None):
        """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks.
        Returns (ms total, last result, wall-clock window)."""
        torch, dist = self.torch, self.dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
            if per_step is not None:
                per_step()
        e1.record()
        torch.cuda.synchronize()
        w1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out, (w0, w1)


def expected_F(cfg_name):
    try:
        v = json.load(open(EXPECTED_F)).get(cfg_name)
        return float(v) if v is not None else None
    except Exception:
        return None


def run_config(env, cfg_name, steps, warmup, args, main_line):
    """Measures one BASELINE configuration; returns the dict of its numbers (rank 0 fills the line with it)."""
    torch, dist = env.torch, env.dist
    from gparml_b200 import _lib
    from gparml_b200.engine import ShardContext
    from gparml_b200.synthetic import split_rows
    rank, world = env.rank, env.world
    g = block_problem_globals(cfg_name)
    N, M, Q, D, fixed = g["N"], g["M"], g["Q"], g["D"], g["fixed_embeddings"]
    lo, hi = split_rows(N, world)[rank]
    n_loc = hi - lo
    rows = block_problem_rows(cfg_name, lo, hi, with_direction=not fixed)

    ctx = ShardContext(M, Q, D, N, device=env.local_rank, fixed_embeddings=fixed, fp32_map=args.fp32)
    ctx.use_torch_stream()
    # pinned host copies of the shard (e2e) and of the per-point gradient
    Yp = torch.from_numpy(rows["Y"]).pin_memory()
    MUp = torch.from_numpy(rows["X_mu"]).pin_memory()
    Sp = torch.from_numpy(rows["X_S"]).pin_memory()
    GLp = torch.empty((2, n_loc, Q), dtype=torch.float64).pin_memory()
    ctx.upload_shard_ptrs(Yp.data_ptr(), MUp.data_ptr(), Sp.data_ptr(), n_loc)
    if not fixed:
        ctx.upload(_lib.A_GRAD_D, rows["d"])
    Z, sf2, alpha, beta = g["Z"], g["sf2"], g["alpha"], g["beta"]
    stats_view = ctx.stats_torch_view() if world > 1 else None
    step_size = 0.0 if fixed else 1e-4

    def evaluation():
        ctx.set_globals(Z, sf2, alpha, beta)
        ctx.set_step(step_size)
        ctx.statistics_launch()          # no host wait: the status word travels with global_step_end
        if world > 1:
            dist.all_reduce(stats_view, op=dist.ReduceOp.SUM)
        if fixed:
            return ctx.global_step()
        ctx.global_step_begin()          # F and the global gradients finish on a side stream ...
        ctx.embedding_grads()            # ... next to the embeddings map, which only needs dF/dPsi1Y, dF/dPsi2
        return ctx.global_step_end()     # F, gradient on the host

    def evaluation_e2e():
        # host buffers in, host buffers out: the Y upload overlaps prep_points + psi2_stats and the
        # gradient download overlaps embed_grads (chunked) inside the library
        ctx.upload_shard_ptrs(Yp.data_ptr(), MUp.data_ptr(), Sp.data_ptr(), n_loc)
        ctx.set_globals(Z, sf2, alpha, beta)
        ctx.set_step(step_size)
        ctx.statistics_launch()
        if world > 1:
            dist.all_reduce(stats_view, op=dist.ReduceOp.SUM)
        if fixed:
            return ctx.global_step()
        ctx.global_step_begin()
        ctx.embedding_grads_into(GLp.data_ptr(), chunks=args.e2e_chunks)
        return ctx.global_step_end()

    for _ in range(warmup):
        evaluation()
    ctx.enable_timing(True)
    launches0 = ctx.launch_count
    phases = {}

    def collect():
        for kk, v in ctx.phase_times_ms().items():
            phases.setdefault(kk, []).append(v)
    ms_total, (F, gflat), win = env.timed(evaluation, steps, per_step=collect)
    launches = ctx.launch_count - launches0
    ctx.enable_timing(False)
    ms_per_step = ms_total / steps
    out = {"config": workload_config(cfg_name, world), "value": 1e3 / ms_per_step, "unit": UNIT, "ms_per_step": ms_per_step,
           "steps": steps, "warmup": warmup, "gpu_launches": launches, "F": F}
    out["clocks"] = env.sampler.window(*win) if env.sampler else None

    # the packed all-reduce, timed alone (device events around `steps` back-to-back all-reduces)
    if world > 1:
        for _ in range(3):
            dist.all_reduce(stats_view, op=dist.ReduceOp.SUM)
        ms_ar, _, _ = env.timed(lambda: dist.all_reduce(stats_view, op=dist.ReduceOp.SUM), 10)
        out["allreduce"] = {"bytes": int(ctx.stats_count) * 8, "ms": ms_ar / 10}
        evaluation()                     # the buffer was summed repeatedly: restore a consistent state

    # F against the committed value of this configuration (the data does not depend on the number of ranks)
    want = expected_F(cfg_name)
    if want is not None and not args.fp32:
        rel = abs(F - want) / abs(want)
        out["F_check"] = {"expected": want, "rel_err": rel, "ok": bool(rel <= 1e-11)}
        if not rel <= 1e-11:
            raise SystemExit("bench.py: %s F = %.17g differs from the committed %.17g (rel %.2e) at %d rank(s)"
                             % (cfg_name, F, want, rel, world))

    if main_line and not args.no_e2e:
        for _ in range(3):
            evaluation_e2e()             # the download sizes its point ranges from the previous call's timings
        ms_e2e, _, _ = env.timed(evaluation_e2e, steps)
        # the phases of the end-to-end step from a few more, instrumented steps (outside the timed region)
        ctx.enable_timing(True)
        ph_e2e = {}
        for _ in range(5):
            evaluation_e2e()
            for kk, v in ctx.phase_times_ms().items():
                ph_e2e.setdefault(kk, []).append(v)
        ctx.enable_timing(False)
        h2d = n_loc * (D + 2 * Q) * 8 + (M * Q + Q + 2) * 8
        d2h = (0 if fixed else 2 * n_loc * Q * 8) + (1 + M * Q + Q + 2) * 8
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device=env.dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        out["e2e"] = {"value": 1e3 / (ms_e2e / steps), "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()),
                      "d2h_bytes_per_step": int(tot[1].item()),
                      "api": "upload_shard(Y, X_mu, X_S) from pinned host memory + evaluation + grad_latest back to pinned host, every step",
                      # same phases as phase_ms_median, here including the waits for the row-range uploads / chunked downloads
                      "phase_ms_median": {kk: float(np.median(v)) for kk, v in ph_e2e.items()}}
        # what the host link gives each rank while ALL ranks copy at once (the copies the e2e step hides behind kernels)
        dY = torch.empty_like(Yp, device=env.dev)
        dG = torch.empty((2, n_loc, Q), dtype=torch.float64, device=env.dev)
        for _ in range(2):
            dY.copy_(Yp, non_blocking=True)
            GLp.copy_(dG, non_blocking=True)
        ms_h2d, _, _ = env.timed(lambda: dY.copy_(Yp, non_blocking=True), 5)
        ms_d2h, _, _ = env.timed(lambda: GLp.copy_(dG, non_blocking=True), 5)
        out["e2e"]["host_link_all_ranks_copying"] = {
            "h2d_gbs_per_rank": Yp.numel() * 8 / (ms_h2d / 5 * 1e-3) / 1e9,
            "d2h_gbs_per_rank": GLp.numel() * 8 / (ms_d2h / 5 * 1e-3) / 1e9,
            "note": "slowest rank (max over ranks of the copy time), pinned host memory, %d-rank concurrent copies" % world}
        del dY, dG

    # ---- rooflines of this rank's kernels, live CUDA-event durations ------------------------------------------
    P = M * (M + 1) // 2
    dfma = ctx.measure_dfma_peak()                       # lane-ops/s, pure DFMA probe on this GPU
    med = {kk: float(np.median(v)) for kk, v in phases.items()}
    w_psi2 = n_loc * P * (6 * Q + E_EXP + 2)
    w_emb = n_loc * (P * (6 * Q + E_EXP + 3) + M * (6 * Q + E_EXP + 3 + 2 * D))
    t_psi2 = med.get("psi2_stats", 0.0) * 1e-3
    ach = 2.0 * w_psi2 / t_psi2 / 1e12 if t_psi2 > 0 else 0.0
    peak = 2.0 * dfma / 1e12
    # "bound": the contract's vocabulary is hbm | tensor; this kernel is compute-bound on the FP64 pipe, which on
    # B200 is also where the FP64 tensor-core instruction executes (same 37.2 TFLOP/s peak, tools/micro/dmma_probe.cu)
    k2x = Q <= 10 and not args.fp32                      # psi2x_stats: exponent from the accumulated u, 5Q + 10 executed
    x_psi2 = n_loc * P * ((5 if k2x else 6) * Q + 10)
    k2_name = "psi2x_stats_kernel" if k2x else "psi2_stats_kernel"
    roofline = {"kernel": "%s<Q=%d>" % (k2_name, Q), "bound": "tensor", "bound_detail": "fp64_pipe (DFMA; DMMA shares it)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak if peak > 0 else None, "traffic": None,
                "peak_source": "pure-DFMA probe kernel measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "peak_nominal": 148 * 64 * 2 * 1.965e9 / 1e12,
                "algorithmic_ops_per_launch": w_psi2, "launch_ms": med.get("psi2_stats"),
                "ops_rule": "FP64-pipe lane-ops, FMA=1, exp=18: n_local * P * (6Q+20), TFLOP/s = 2*ops/t (SURVEY.md 8d)",
                # what the kernel actually issues (table-driven exp = 8 instructions; psi2x_stats 5 instead of 6 per
                # latent dimension): the FP64-pipe busy fraction.  frac counts the survey's 6Q+20 and can exceed it
                "executed_ops_per_launch": x_psi2,
                "executed_frac": (x_psi2 / t_psi2 / dfma) if t_psi2 > 0 and dfma > 0 else None,
                "note": "frac uses the algorithmic count of SURVEY.md 8d; the kernel issues %dQ+10 FP64 instructions per "
                        "point-pair, so executed_frac is the pipe-busy fraction" % (5 if k2x else 6)}
    if not fixed and med.get("embed_grads", 0) > 0:
        t_emb = med["embed_grads"] * 1e-3
        k5m = 5 <= Q <= 10 and not args.fp32                  # embed_psi2m: both products on the FP64 tensor-core instruction
        if k5m:      # per (point, pair): (ceil(2Q/4) + 2 ceil((2Q+1)/8)) MMAs of 256 FMAs per 64 items + 6 (exp)
            per_pair = ((2 * Q + 3) // 4 + 2 * ((2 * Q + 8) // 8)) * 4 + 6
        else:
            per_pair = 4 * Q + 9                                # embed_psi2x; exp = 7 instructions (256-entry table)
        x_emb = n_loc * (P * per_pair + M * (6 * Q + 10 + 2 * D))
        roofline["embed_grads"] = {"achieved": 2.0 * w_emb / t_emb / 1e12, "frac": (2.0 * w_emb / t_emb / 1e12) / peak,
                                   "launch_ms": med["embed_grads"], "algorithmic_ops_per_launch": w_emb,
                                   "executed_ops_per_launch": x_emb, "executed_frac": x_emb / t_emb / dfma if dfma > 0 else None,
                                   "kernel": "embed_psi2m_kernel (FP64 MMA)" if k5m else "embed_psi2x_kernel",
                                   "note": "algorithmic count of SURVEY.md 8d (6Q+21 per point-pair); the expanded-basis kernels "
                                           "execute %d FP64-pipe lane-ops per point-pair (MMA lanes included), so frac can exceed "
                                           "the pipe-busy fraction (executed_frac)" % per_pair}
    if args.fp32:
        roofline["note"] = "fp32 map kernels selected: the FP64-pipe roofline above does not describe them"
    total_ops = algorithmic_ops(N, M, Q, D, fixed)
    roofline["whole_evaluation_frac"] = (2.0 * total_ops / world / (ms_per_step * 1e-3) / 1e12) / peak
    # DRAM traffic from the committed ncu --set full captures of the same workload
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", args.traffic_file))).get(cfg_name, {})
    except Exception:
        pass

    def dram(kernel):
        ent = traffic.get(kernel)
        if ent and int(ent["n_local"]) == n_loc:
            return ent["dram_bytes"], ent.get("source")
        return None, None
    roofline["traffic"], src = dram(k2_name)
    if src:
        roofline["traffic_source"] = src
    # the HBM-bound streaming kernels: algorithmic bytes against the measured copy bandwidth
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        pass
    R = (3 * Q + 2) & ~1
    prep_bytes = n_loc * 8 * ((2 * Q if fixed else 4 * Q) + 2 * R + 2 * Q + ((4 * Q + 2) if k2x else 0))
    if med.get("prep_points", 0) > 0:
        gbs = prep_bytes / (med["prep_points"] * 1e-3) / 1e9
        roofline["prep_points_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                       "peak_source": hbm_src, "algorithmic_bytes_per_launch": prep_bytes,
                                       "launch_ms": med["prep_points"], "traffic": dram("prep_points_kernel")[0]}
    # K6, the optimiser's local-state passes (scg_adapted_local_MapReduce.py:29-243 on device-resident vectors): pure streams
    if not fixed and main_line:
        vec_bytes = 2 * n_loc * Q * 8
        for _ in range(2):
            ctx.scg_update_d(0.5)
            ctx.scg_update_grad_old()
        ms_axpy, _, _ = env.timed(lambda: ctx.scg_update_d(0.5), 10)          # d = g d - new: 2 reads + 1 write
        ms_copy, _, _ = env.timed(lambda: ctx.scg_update_grad_old(), 10)      # old = new: 1 read + 1 write
        roofline["scg_local_state_hbm"] = {
            "bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src,
            # at several GPUs a rank's vectors fit the 126 MB L2: the pass is then L2- not HBM-bound and frac can exceed 1
            "l2_resident": bool(3 * vec_bytes < 100e6),
            "update_d": {"achieved": 3 * vec_bytes / (ms_axpy / 10 * 1e-3) / 1e9, "frac": 3 * vec_bytes / (ms_axpy / 10 * 1e-3) / 1e9 / hbm_peak,
                         "algorithmic_bytes_per_launch": 3 * vec_bytes, "launch_ms": ms_axpy / 10, "traffic": dram("scg_update_kernel")[0]},
            "update_grad_old": {"achieved": 2 * vec_bytes / (ms_copy / 10 * 1e-3) / 1e9, "frac": 2 * vec_bytes / (ms_copy / 10 * 1e-3) / 1e9 / hbm_peak,
                                "algorithmic_bytes_per_launch": 2 * vec_bytes, "launch_ms": ms_copy / 10}}
    out["roofline"] = roofline
    out["phase_ms_median"] = med

    # ---- the same evaluation through the reference's interface (parallel_GPLVM protocol on b200_MapReduce) -----
    if main_line and not args.no_through_driver and not args.fp32:
        ctx.synchronize()
        try:
            out["through_driver"] = through_driver(env, cfg_name, g, rows, steps, warmup, step_size, F)
        except SystemExit:
            raise
        except Exception as e:      # e.g. no room for the shard files: the main line must survive (all ranks fail alike)
            out["through_driver"] = {"value": None, "unit": UNIT, "error": repr(e)[:300]}
    ctx.close()
    del Yp, MUp, Sp, GLp, rows
    return out


def through_driver(env, cfg_name, g, rows, steps, warmup, step_size, F_raw):
    """Times ``parallel_GPLVM.likelihood_and_gradient(x, i, step)`` -- the callback the reference's optimiser calls
    (parallel_GPLVM.py:222-279) -- on the b200_MapReduce backend: one input file per rank (``.npy`` shards), the
    ``--load`` path for the initial state (the 'f' checkpoint files are written below), no per-evaluation files."""
    from gparml_b200 import b200_MapReduce, parallel_GPLVM as drv
    rank, world = env.rank, env.world
    need = sum(int(v.nbytes) for v in rows.values() if v is not None) * world * 1.25 + (64 << 20)
    base = None
    for cand in ("/dev/shm", tempfile.gettempdir()):       # shared-memory file system if it has the room, else the temp dir
        try:
            if os.path.isdir(cand) and shutil.disk_usage(cand).free > need:
                base = cand
                break
        except OSError:
            pass
    work = [tempfile.mkdtemp(prefix="gparml_bench_", dir=base) if rank == 0 else None]
    if world > 1:
        env.dist.broadcast_object_list(work, src=0, device=env.dev)
    work = work[0]
    dirs = {d: os.path.join(work, d) for d in ("input", "embeddings", "statistics", "tmp")}
    try:
        if rank == 0:
            for d in dirs.values():
                os.makedirs(d)
            for key, val in (("Z", g["Z"]), ("sf2", np.array([[g["sf2"]]])), ("alpha", g["alpha"].reshape(1, -1)),
                             ("beta", np.array([[g["beta"]]]))):
                np.save(os.path.join(dirs["statistics"], "global_statistics_%s_f.npy" % key), val)
        if world > 1:
            env.dist.barrier()
        name = "shard_%03d.npy" % rank
        ok, why = 1, ""
        try:
            np.save(os.path.join(dirs["input"], name), rows["Y"])
            np.save(os.path.join(dirs["embeddings"], name + ".embedding.npy"), rows["X_mu"])
            np.save(os.path.join(dirs["embeddings"], name + ".variance.npy"), rows["X_S"])
            if "d" in rows:
                np.save(os.path.join(dirs["embeddings"], name + ".grad_d.npy"), rows["d"])
        except Exception as e:
            ok, why = 0, repr(e)
        if world > 1:           # all ranks go on, or none (a rank that failed alone would leave the others in a collective)
            flag = env.torch.tensor([ok], dtype=env.torch.int32, device=env.dev)
            env.dist.all_reduce(flag, op=env.dist.ReduceOp.MIN)
            ok = int(flag.item())
        if not ok:
            raise RuntimeError("could not stage the shard files in %s: %s" % (work, why or "another rank failed"))
        # b200_stream='torch': the shard contexts work on torch's current stream, the one the CUDA events of
        # Env.timed are recorded on (on their own streams the last step's embeddings map would fall outside e0..e1)
        opts = drv.default_options(M=g["M"], Q=g["Q"], D=g["D"], load=True, fixed_embeddings=g["fixed_embeddings"],
                                   b200_write_files=False, b200_stream="torch", display=False, **dirs)
        opts = b200_MapReduce.init(opts)
        opts, gs = drv.init_statistics(b200_MapReduce, opts)
        x0 = drv.flatten_global_statistics(opts, gs)
        x0 = np.array([drv.sp.transform_back(b, x) for b, x in zip(opts["flat_global_statistics_bounds"], x0)])
        drv.options, drv.map_reduce = opts, b200_MapReduce
        it = [0]

        def call():
            it[0] += 1
            return drv.likelihood_and_gradient(x0, it[0], step_size)
        for _ in range(max(warmup, 1)):
            f, grad = call()
        ms, (f, grad), _ = env.timed(call, steps)
        rel = abs(-f - F_raw) / abs(F_raw)
        if not rel <= 1e-12:
            raise SystemExit("bench.py: F through the driver (%.17g) differs from the C-ABI loop (%.17g)" % (-f, F_raw))
        return {"value": 1e3 / (ms / steps), "unit": UNIT, "ms_per_step": ms / steps,
                "api": "gparml_b200.parallel_GPLVM.likelihood_and_gradient(x, i, step_size) on b200_MapReduce "
                       "(parallel_GPLVM.py:222-279 protocol, one input shard per rank, b200_write_files=False)",
                "F_matches_c_abi_loop": True}
    finally:
        b200_MapReduce.close()
        if world > 1:
            env.dist.barrier()
        if rank == 0:
            shutil.rmtree(work, ignore_errors=True)


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--other-configs", default="c5,c4,c2", help="comma-separated BASELINE configs measured after the main one "
                                                                "('' or 'none' = none); only with the default main config c3")
    ap.add_argument("--other-steps", type=int, default=3)
    ap.add_argument("--cpu-points", type=int, default=0, help="points per CPU worker in the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-through-driver", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=3, help="point ranges of the overlapped gradient download (1..8)")
    ap.add_argument("--traffic-file", default="ncu_traffic_r02.json")
    ap.add_argument("--fp32", action="store_true", help="opt-in fp32 map path (psi2_stats / embed_grads in fp32, fp64 sums); "
                                                        "not the headline configuration")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    env = Env()
    rank, world = env.rank, env.world
    res = run_config(env, args.config, args.steps, args.warmup, args, main_line=True)
    others = []
    if args.config == "c3" and not args.fp32:
        names = [c.strip().strip('"\'') for c in args.other_configs.split(",")]
        for name in [c for c in names if c and c.lower() != "none"]:
            o = run_config(env, name, args.other_steps, 3, args, main_line=False)
            r = o["roofline"]
            others.append({"config": o["config"], "value": o["value"], "unit": UNIT, "ms_per_step": o["ms_per_step"],
                           "steps": o["steps"], "warmup": o["warmup"], "clocks": o["clocks"], "F": o["F"],
                           "F_check": o.get("F_check"), "allreduce": o.get("allreduce"),
                           "psi2_stats": {"frac": r["frac"], "executed_frac": r["executed_frac"], "launch_ms": r["launch_ms"]},
                           "embed_grads": ({"frac": r["embed_grads"]["frac"], "executed_frac": r["embed_grads"]["executed_frac"],
                                            "launch_ms": r["embed_grads"]["launch_ms"]} if "embed_grads" in r else None),
                           "whole_evaluation_frac": r["whole_evaluation_frac"], "phase_ms_median": o["phase_ms_median"]})
    if env.sampler:
        env.sampler.stop()

    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 maps, f64 sums and master step (opt-in)" if args.fp32 else "f64",
        "data": "synthetic", "config": res["config"], "clocks": res["clocks"], "e2e": res.get("e2e"),
        "through_driver": res.get("through_driver"),
        "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "phase_ms_median": res["phase_ms_median"],
        "F": res["F"], "F_check": res.get("F_check"), "allreduce": res.get("allreduce"), "other_configs": others,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            pts = args.cpu_points or default_cpu_points(args.config, "baseline")
            cb = CpuArm(args.config, cores, pts)
            cb.step()                      # warm-up: worker imports, first-touch (the reference arm warms up too)
            tm, tg = cb.step()
            cb.close()
            line["cpu_baseline"] = {"value": cb.evals_per_s(tm, tg), "unit": UNIT, "cores": cores, "kind": cb.kind,
                                    "sample": cb.describe(), "sample_wall_s": tm + tg}
        except Exception as e:  # the baseline must never take the GPU line down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
