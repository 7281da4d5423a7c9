#!/usr/bin/env python
"""Kernel-variant tuning harness (development tool, not part of the product path).

  python tools/tune.py build                 # here (CPU): compile the variants listed in VARIANTS
  python tools/tune.py run [N]               # on the GPU box: time every variant, print a table
  python tools/tune.py one <lib.so> [N]      # internal: time one variant library

Variants are the production sources compiled with different -D tuning macros; each becomes
gparml_b200/variants/lib_<name>.so, selected at run time through GPARML_B200_LIB.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "gparml_b200", "csrc")
VDIR = os.path.join(ROOT, "gparml_b200", "variants")
NVCC = "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
from gparml_b200.build import SOURCES as ALL  # noqa: E402

# name -> {source: [defines]}
VARIANTS = {
    "base": {},
    # e.g. "x_tn64": {"psi2.cu": ["PSI2X_TN=64"]},  "k5_dfma": {"embed.cu": ["EMB_NO_MMA"]},  "k5m_exp8": {"embed_m.cu": ["EMBM_EXP12=0"]},
    #      "p1m_tp32": {"psi1_mma.cu": ["P1M_TP=32"]},  a leading "-" passes an nvcc flag instead of a -D macro
}


def build(names=None):
    os.makedirs(VDIR, exist_ok=True)
    from gparml_b200 import build as b
    b.build()
    base_objs = {s: os.path.join(b.OBJ, s.replace(".cu", ".o")) for s in ALL}
    for name, spec in VARIANTS.items():
        if names and name not in names:
            continue
        objs = dict(base_objs)
        log = []
        for src, defs in spec.items():
            o = os.path.join(VDIR, "%s_%s.o" % (name, src.replace(".cu", "")))
            cmd = [NVCC] + FLAGS + ["-Xptxas", "-v"] + [d if d.startswith("-") else "-D" + d for d in defs] + ["-c", os.path.join(CSRC, src), "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                print("variant %s failed:\n%s" % (name, r.stderr[-3000:]))
                objs = None
                break
            objs[src] = o
            # registers of the Q=10 instantiation
            lines = r.stderr.splitlines()
            for i, ln in enumerate(lines):
                if "ILi10E" in ln and "Compiling entry" in ln:
                    log.append(" ".join(x.strip() for x in lines[i + 1:i + 3]))
        if objs is None:
            continue
        lib = os.path.join(VDIR, "lib_%s.so" % name)
        subprocess.check_call([NVCC, "-shared", "-o", lib] + list(objs.values()) + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
        print(name, "|", " || ".join(log))
    for f in os.listdir(VDIR):
        if f.endswith(".o"):
            os.remove(os.path.join(VDIR, f))


def one(lib, n):
    os.environ["GPARML_B200_LIB"] = lib
    import numpy as np
    from gparml_b200 import _lib
    from gparml_b200.engine import ShardContext
    from gparml_b200.synthetic import CONFIGS, make_problem
    k = CONFIGS[os.environ.get("GPARML_TUNE_CFG", "c3")]
    fixed = bool(k.get("fixed_embeddings"))
    p = make_problem(n, k["M"], k["Q"], k["D"], seed=3, with_direction=not fixed, fixed_embeddings=fixed)
    c = ShardContext(k["M"], k["Q"], k["D"], n, fp32_map=os.environ.get("GPARML_TUNE_FP32") == "1", fixed_embeddings=fixed)
    c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
    c.enable_timing(True)
    acc = {}
    for it in range(5):
        c.set_globals(p["Z"], p["sf2"], p["alpha"], p["beta"])
        c.statistics()
        if fixed:
            F, g = c.global_step()
        else:                              # like bench.py: the tail of the master step runs next to the embeddings map
            c.global_step_begin()
            c.embedding_grads()
            F, g = c.global_step_end()
        t = c.phase_times_ms()
        if it >= 2:
            for kk, v in t.items():
                acc.setdefault(kk, []).append(v)
    out = {kk: float(np.median(v)) for kk, v in acc.items()}
    out["F"] = F
    out["chk"] = 0.0 if fixed else float(np.abs(c.grad_latest()).sum())
    print(json.dumps(out))


def run(n):
    rows = []
    for f in sorted(os.listdir(VDIR)):
        if not f.endswith(".so"):
            continue
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", os.path.join(VDIR, f), str(n)],
                           capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            print(f, "FAILED", r.stderr[-500:])
            continue
        rows.append((f, d))
        print("%-28s psi2 %7.3f  embed %7.3f  psi1 %6.3f  prep %5.3f  glob %5.3f   F=%.10g chk=%.10g" % (
            f, d["psi2_stats"], d["embed_grads"], d["psi1_stats"], d["prep_points"], d["global_step"], d["F"], d["chk"]))
    return rows


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "build":
        build(sys.argv[2:] or None)
    elif cmd == "run":
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 250000)
    elif cmd == "one":
        one(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 250000)
