#!/usr/bin/env python
"""Benchmark of the map-reduce variational-bound hot path (BASELINE.json metric:
ELBO+gradient evaluations per second at N=1M, M=100, Q=10, D=10 on 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one full evaluation: globals host->device, prep_points,
psi1_stats, psi2_stats, [NCCL all-reduce of the packed sums], global_step (F and the global
gradient back on the host), embed_grads.  N > 1 shards the points over the ranks
(strong scaling: the problem size is fixed by the metric).

``value``  : device-resident shard (Y, X_mu, X_S uploaded once before the timed region).
``e2e``    : the same evaluation through the host-buffer API every step: the shard is copied
             from pinned host memory (what ``partial_terms.set_data`` receives in the
             reference's mappers, local_MapReduce.py:197-224) and the per-point gradients
             are copied back (the ``.grad_latest.npy`` the reference writes, :359-360).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line (rank 0): NCCL's own banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gparml_b200.synthetic import CONFIGS, make_problem, split_rows  # noqa: E402

E_EXP = 18  # FP64-pipe instructions of exp()
METRIC = "ELBO+grad evals/sec"
UNIT = "evals/s"


def algorithmic_ops(N, M, Q, D, fixed):
    """FP64-pipe lane-ops of one evaluation over N points."""
    P = M * (M + 1) // 2
    if fixed:
        return N * (P * (6 * Q + E_EXP + 2) + M * (6 * Q + E_EXP + 1 + D + 2 * Q * D))
    return N * (P * (12 * Q + 2 * E_EXP + 5) + M * (12 * Q + 2 * E_EXP + 4 + 3 * D + 2 * Q * D))


# ----------------------------------------------------------------------------------------
# CPU baseline: the numpy oracle (a port of the reference's per-point numpy loops) hosted in a
# multiprocessing.Pool with one shard per worker, as local_MapReduce.py:134 does.
# ----------------------------------------------------------------------------------------
def _cpu_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"


def _cpu_stats_worker(a):
    from oracle import gparml_oracle as O
    sh, Z, sf2, alpha, fixed = a
    t = time.time()
    mu, S, _ = O.effective_embedding(sh["X_mu"], sh["X_S"], None, 0.0, fixed)
    st = O.shard_statistics_chunked(sh["Y"], mu, S, Z, sf2, alpha, chunk=256)
    return st, time.time() - t


def _cpu_embed_worker(a):
    from oracle import gparml_oracle as O
    sh, Z, sf2, alpha, G1, G2 = a
    t = time.time()
    mu, S, sraw = O.effective_embedding(sh["X_mu"], sh["X_S"], None, 0.0, False)
    gm, gs = O.embedding_grads(sh["Y"], mu, S, Z, sf2, alpha, G1, G2)
    g = -np.array([gm, gs * O.softplus_grad(sraw)])
    return float(np.abs(g).sum()), time.time() - t


class CpuBaseline(object):
    def __init__(self, cfg_name, cores, pts_per_worker):
        import multiprocessing
        from oracle import gparml_oracle as O  # noqa: F401  (the one place bench.py may use oracle/)
        k = CONFIGS[cfg_name]
        self.k, self.cores, self.pts = k, cores, pts_per_worker
        n = cores * pts_per_worker
        p = make_problem(n, k["M"], k["Q"], k["D"], seed=int(cfg_name[1]), fixed_embeddings=k["fixed_embeddings"])
        self.p = p
        self.shards = [dict(Y=p["Y"][lo:hi], X_mu=p["X_mu"][lo:hi], X_S=p["X_S"][lo:hi]) for lo, hi in split_rows(n, cores)]
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
        from oracle import gparml_oracle as O
        p, k = self.p, self.k
        t0 = time.time()
        res = self.pool.map(_cpu_stats_worker, [(s, p["Z"], p["sf2"], p["alpha"], k["fixed_embeddings"]) for s in self.shards])
        t_map = time.time() - t0
        t1 = time.time()
        stats = O.reduce_statistics([r[0] for r in res])
        g = O.global_step(stats, p["Z"], p["sf2"], p["alpha"], p["beta"], len(p["Y"]))
        t_glob = time.time() - t1
        if not k["fixed_embeddings"]:
            t2 = time.time()
            self.pool.map(_cpu_embed_worker, [(s, p["Z"], p["sf2"], p["alpha"], g["dF_dsum_exp_K_miY"],
                                               g["dF_dsum_exp_K_mi_K_im"]) for s in self.shards])
            t_map += time.time() - t2
        return t_map, t_glob

    def close(self):
        self.pool.close()
        self.pool.join()

    def evals_per_s(self, t_map, t_glob):
        """Linear extrapolation of the maps to the full N (exactly linear: one Python iteration
        per point, partial_terms.py:46,200,278,383,416) plus the master step once."""
        full = t_map * (self.k["N"] / float(self.cores * self.pts)) + t_glob
        return 1.0 / full

    def describe(self):
        return ("numpy port of the reference maps in a %d-process Pool, %d points per worker (%d of %d points), "
                "one evaluation, map time extrapolated linearly to N; CSV parsing and .npy transport excluded"
                % (self.cores, self.pts, self.cores * self.pts, self.k["N"]))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    k = CONFIGS[args.config]
    pts = args.cpu_points or (8 if k["M"] >= 400 else 256)
    cb = CpuBaseline(args.config, cores, pts)
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
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": cb.describe()},
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

    def stop(self, t0=None, t1=None):
        """Summary of the samples received in the wall-clock window [t0, t1] (the timed regions)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [(t, ln) for t, ln in self.lines if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.15)]
        if not rows:
            rows = self.lines[-2:]
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
    ap.add_argument("--n", type=int, default=0, help="override the total number of points (debug only; marks the line invalid)")
    ap.add_argument("--cpu-points", type=int, default=0, help="points per CPU worker in the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="point ranges of the overlapped gradient download (1..8)")
    ap.add_argument("--fp32", action="store_true", help="opt-in fp32 map path (psi2_stats / embed_grads in fp32, fp64 sums); "
                                                        "not the headline configuration")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from gparml_b200 import _lib
    from gparml_b200.engine import ShardContext

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- gparml_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry
        # exactly one JSON line, so file descriptor 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved_fd = os.dup(1)
        try:
            os.dup2(2, 1)
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    k = dict(CONFIGS[args.config])
    if args.n:
        k["N"] = args.n
    N, M, Q, D, fixed = k["N"], k["M"], k["Q"], k["D"], k["fixed_embeddings"]
    p = make_problem(N, M, Q, D, seed=int(args.config[1]), fixed_embeddings=fixed, with_direction=not fixed)
    lo, hi = split_rows(N, world)[rank]
    n_loc = hi - lo

    ctx = ShardContext(M, Q, D, N, device=local_rank, fixed_embeddings=fixed, fp32_map=args.fp32)
    ctx.use_torch_stream()
    # pinned host copies of the shard (e2e) and of the per-point gradient
    Yp = torch.from_numpy(np.ascontiguousarray(p["Y"][lo:hi])).pin_memory()
    MUp = torch.from_numpy(np.ascontiguousarray(p["X_mu"][lo:hi])).pin_memory()
    Sp = torch.from_numpy(np.ascontiguousarray(p["X_S"][lo:hi])).pin_memory()
    GLp = torch.empty((2, n_loc, Q), dtype=torch.float64).pin_memory()
    ctx.upload_shard_ptrs(Yp.data_ptr(), MUp.data_ptr(), Sp.data_ptr(), n_loc)
    if not fixed:
        ctx.upload(_lib.A_GRAD_D, p["d"][:, lo:hi])
    Z, sf2, alpha, beta = p["Z"], p["sf2"], p["alpha"], p["beta"]
    del p
    stats_view = ctx.stats_torch_view() if world > 1 else None
    step_size = 0.0 if fixed else 1e-4

    def evaluation():
        ctx.set_globals(Z, sf2, alpha, beta)
        ctx.set_step(step_size)
        ctx.statistics()
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
        ctx.statistics()
        if world > 1:
            dist.all_reduce(stats_view, op=dist.ReduceOp.SUM)
        if fixed:
            return ctx.global_step()
        ctx.global_step_begin()
        ctx.embedding_grads_into(GLp.data_ptr(), chunks=args.e2e_chunks)
        return ctx.global_step_end()

    def timed(fn, steps, collect_phases=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = {}
        e0.record()
        for _ in range(steps):
            out = fn()
            if collect_phases:
                for kk, v in ctx.phase_times_ms().items():
                    phases.setdefault(kk, []).append(v)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out, phases

    uuid = ""
    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        pass
    sampler = ClockSampler(uuid) if rank == 0 else None
    if sampler:
        sampler.start()          # started before the warm-up so that it is sampling when the timed region begins
    for _ in range(args.warmup):
        evaluation()
    ctx.enable_timing(True)
    launches0 = ctx.launch_count
    t_load0 = time.time()
    ms_total, (F, g), phases = timed(evaluation, args.steps, collect_phases=True)
    launches = ctx.launch_count - launches0
    ctx.enable_timing(False)
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    e2e = None
    if not args.no_e2e:
        evaluation_e2e()
        ms_e2e, _, _ = timed(evaluation_e2e, args.steps)
        h2d = n_loc * (D + 2 * Q) * 8 + (M * Q + Q + 2) * 8
        d2h = (0 if fixed else 2 * n_loc * Q * 8) + (1 + M * Q + Q + 2) * 8
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        e2e = {"value": 1e3 / (ms_e2e / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()),
               "d2h_bytes_per_step": int(tot[1].item()),
               "api": "upload_shard(Y, X_mu, X_S) from pinned host memory + evaluation + grad_latest back to pinned host, every step"}

    clocks = sampler.stop(t_load0, time.time()) if sampler else None   # samples during the two timed loops

    # roofline of the dominant kernel (psi2_stats) on this rank, live CUDA-event durations
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
    roofline = {"kernel": "psi2_stats_kernel<Q=%d>" % Q, "bound": "tensor", "bound_detail": "fp64_pipe (DFMA; DMMA shares it)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak if peak > 0 else None, "traffic": None,
                "peak_source": "pure-DFMA probe kernel measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "peak_nominal": 148 * 64 * 2 * 1.965e9 / 1e12,
                "algorithmic_ops_per_launch": w_psi2, "launch_ms": med.get("psi2_stats"),
                "ops_rule": "FP64-pipe lane-ops, FMA=1, exp=18: n_local * P * (6Q+20), TFLOP/s = 2*ops/t (SURVEY.md 8d)",
                # what the kernel actually issues (table-driven exp = 9): the FP64-pipe busy fraction
                "executed_ops_per_launch": n_loc * P * (6 * Q + 11),
                "executed_frac": (n_loc * P * (6 * Q + 11) / t_psi2 / dfma) if t_psi2 > 0 and dfma > 0 else None}
    if not fixed and med.get("embed_grads", 0) > 0:
        t_emb = med["embed_grads"] * 1e-3
        x_emb = n_loc * (P * (4 * Q + 11) + M * (6 * Q + 12 + 2 * D))
        roofline["embed_grads"] = {"achieved": 2.0 * w_emb / t_emb / 1e12, "frac": (2.0 * w_emb / t_emb / 1e12) / peak,
                                   "launch_ms": med["embed_grads"], "algorithmic_ops_per_launch": w_emb,
                                   "executed_ops_per_launch": x_emb, "executed_frac": x_emb / t_emb / dfma if dfma > 0 else None,
                                   "note": "algorithmic count of SURVEY.md 8d (6Q+21 per point-pair); the expanded-basis kernel "
                                           "issues 4Q+11, so frac can exceed the pipe-busy fraction (executed_frac)"}
    if args.fp32:
        roofline["note"] = "fp32 map kernels selected: the FP64-pipe roofline above does not describe them"
    total_ops = algorithmic_ops(N, M, Q, D, fixed)
    roofline["whole_evaluation_frac"] = (2.0 * total_ops / world / (ms_per_step * 1e-3) / 1e12) / peak
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of the same workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r01c.json")))
        ent = tr.get(args.config, {}).get("psi2_stats_kernel")
        if ent and int(ent["n_local"]) == n_loc:
            roofline["traffic"] = ent["dram_bytes"]
            roofline["traffic_source"] = ent.get("source")
    except Exception:
        pass
    # the HBM-bound streaming kernel (prep_points): algorithmic bytes against the measured copy bandwidth
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        pass
    R = (3 * Q + 2) & ~1
    prep_bytes = n_loc * 8 * ((2 * Q if fixed else 4 * Q) + 2 * R + 2 * Q)
    if med.get("prep_points", 0) > 0:
        gbs = prep_bytes / (med["prep_points"] * 1e-3) / 1e9
        roofline["prep_points_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                       "peak_source": hbm_src, "algorithmic_bytes_per_launch": prep_bytes,
                                       "launch_ms": med["prep_points"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 maps, f64 sums and master step (opt-in)" if args.fp32 else "f64",
        "data": "synthetic", "config": workload_config(args.config, world), "clocks": clocks, "e2e": e2e,
        "gpu_launches": launches, "roofline": roofline, "phase_ms_median": med,
        "F": F,
    }
    if args.n:
        line["config"]["debug_n_override"] = args.n

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            pts = args.cpu_points or (16 if M >= 400 else 1024)
            cb = CpuBaseline(args.config, cores, pts)
            tm, tg = cb.step()
            cb.close()
            line["cpu_baseline"] = {"value": cb.evals_per_s(tm, tg), "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": cb.describe(), "sample_wall_s": tm + tg}
        except Exception as e:  # the baseline must never take the GPU line down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    ctx.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
