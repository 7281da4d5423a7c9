"""In-tree build of the CUDA library: ``python -m gparml_b200.build [-v] [--force]``.

nvcc cross-compiles for sm_100a without a GPU.  The translation units compile in
parallel and link into ``gparml_b200/libgparml_b200.so`` (git-ignored; travels to
the GPU box with the snapshot).
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libgparml_b200.so")
SOURCES = ["capi.cu", "prep.cu", "psi1.cu", "psi1_mma.cu", "psi1_wide.cu", "psi2.cu", "psi2_f32.cu", "embed.cu", "embed_x.cu", "embed_m.cu", "global_step.cu", "global_step_large.cu", "misc.cu", "init.cu"]
HEADERS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".inc"))] + [
    os.path.join(os.path.dirname(HERE), "include", "gparml_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
EXTRA_FLAGS = {}     # per-source extra nvcc flags


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if _newer(obj, [path] + HEADERS):
        return obj, ""
    cmd = [NVCC] + FLAGS + EXTRA_FLAGS.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if not _newer(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
