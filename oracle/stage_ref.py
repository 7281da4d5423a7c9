"""TEST / BENCH INFRASTRUCTURE ONLY -- stage the reference's three hot-path modules for the GPU box.

``/root/reference`` exists only in the build container.  ``bench.py --impl reference`` and the
``cpu_baseline`` leg must time the *reference's own* CPU arithmetic on the GPU box's host cores, so this recipe
copies ``kernels.py``, ``kernel_exp.py`` and ``partial_terms.py`` byte for byte into ``oracle/_ref/`` --
git-ignored (nothing of the reference enters the history) but not gpurun-ignored (it travels with the snapshot
like the built ``.so`` files).  ``oracle/ref_shim.py`` loads them from there when ``/root/reference`` is absent.

    python -m oracle.stage_ref          # also run by __graft_entry__.build() when /root/reference is present
"""
import filecmp
import os
import shutil

SRC = os.environ.get("GPARML_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("kernels.py", "kernel_exp.py", "partial_terms.py")


def stage():
    """Returns the list of staged files, or [] when the reference tree is not present here."""
    if not all(os.path.isfile(os.path.join(SRC, f)) for f in FILES):
        return []
    os.makedirs(DST, exist_ok=True)
    out = []
    for f in FILES:
        a, b = os.path.join(SRC, f), os.path.join(DST, f)
        if not (os.path.isfile(b) and filecmp.cmp(a, b, shallow=False)):
            shutil.copyfile(a, b)
        out.append(b)
    return out


if __name__ == "__main__":
    print("\n".join(stage()) or "reference tree not present at %s: nothing staged" % SRC)
