"""CPU checks of the constants behind the device exp (gparml_b200/csrc/gp_exp.cuh, embed_m.cu): the committed table is
what tools/gen_exp_table.py writes, and a numpy emulation of the three device forms -- same constants, parsed from the
sources, same operation order with fma replaced by mul + add -- stays inside the accuracy DESIGN.md quotes."""
import math
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gparml_b200", "csrc")


def _defines(path, block=None):
    text = open(path).read()
    if block is not None:
        a, b = block
        text = text[text.index(a):text.index(b)]
    out = {}
    for m in re.finditer(r"#define\s+(\w+)\s+\(?(-?[0-9][0-9.eE+\-/ ]*)\)?\s*(?://.*)?$", text, re.M):
        try:
            out[m.group(1)] = float(eval(m.group(2)))
        except Exception:
            pass
    return out


def _table(path):
    vals = []
    for line in open(path):
        if line.lstrip().startswith("//"):
            continue
        vals += [float(v) for v in line.replace(",", " ").split()]
    return np.array(vals)


def _emulate(x, scale, neg_step, log2n, table, coeffs):
    """exp(x) the way the kernels evaluate it: k = round(x scale), r = x - k step, 2^(k >> log2n) table[k & (N-1)] p(r)."""
    k = np.rint(x * scale)
    r = k * neg_step + x
    p = np.full_like(x, coeffs[-1])
    for c in coeffs[-2::-1]:
        p = p * r + c
    ki = k.astype(np.int64)
    return np.ldexp(table[ki & ((1 << log2n) - 1)] * p, (ki >> log2n).astype(np.int64))


def test_committed_table_is_what_the_generator_writes(tmp_path):
    out = str(tmp_path / "table8.inc")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_exp_table.py"), "8", out], stdout=subprocess.DEVNULL)
    assert open(out).read() == open(os.path.join(CSRC, "gp_exp_table8.inc")).read()
    t = _table(os.path.join(CSRC, "gp_exp_table8.inc"))
    assert t.shape == (256,) and t[0] == 1.0
    assert np.max(np.abs(t / np.exp2(np.arange(256) / 256.0) - 1.0)) < 2.3e-16


def test_device_exp_forms_meet_their_stated_accuracy():
    rng = np.random.default_rng(7)
    x = np.concatenate([-rng.uniform(0.0, 60.0, 200000), rng.uniform(0.0, 30.0, 20000), -rng.uniform(60.0, 700.0, 20000)])
    ref = np.exp(x.astype(np.longdouble))
    hdr = os.path.join(CSRC, "gp_exp.cuh")
    # 256 entries, degree-3 near-minimax (the default: K5b, K1, embed_psi2x)
    d8 = _defines(hdr, ("#if GP_EXP_LOG2_TAB == 8", "#elif GP_EXP_LOG2_TAB == 6"))
    t8 = _table(os.path.join(CSRC, "gp_exp_table8.inc"))
    assert abs(d8["GP_EXP_SCALE"] - 256 / math.log(2)) < 1e-12 and abs(d8["GP_EXP_NEG_STEP"] + math.log(2) / 256) < 1e-18
    y8 = _emulate(x, d8["GP_EXP_SCALE"], d8["GP_EXP_NEG_STEP"], 8, t8, [d8["GP_EXP_C0"], d8["GP_EXP_C1"], d8["GP_EXP_C2"], d8["GP_EXP_C3"]])
    e8 = float(np.max(np.abs(y8 / ref - 1)))
    # 64 entries, degree-4 Taylor (psi2.cu)
    d6 = _defines(hdr, ("#elif GP_EXP_LOG2_TAB == 6", "#else\n#error"))
    t6 = np.array([float(v) for v in re.findall(r"[0-9]\.[0-9]+", open(hdr).read()[open(hdr).read().index("#define GP_EXP_TABLE_VALUES"):open(hdr).read().index("static __device__ const double gp_exp_table_const[GP_EXP_TAB] = {GP_EXP_TABLE_VALUES}")])])
    assert t6.shape == (64,)
    y6 = _emulate(x, d6["GP_EXP_SCALE"], d6["GP_EXP_NEG_STEP"], 6, t6, [1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0])
    e6 = float(np.max(np.abs(y6 / ref - 1)))
    # 4096 entries = 256-entry table x 16 sub-steps, degree-2 near-minimax (embed_psi2m)
    src = os.path.join(CSRC, "embed_m.cu")
    dm = _defines(src, ("#if EMBM_EXP12", "#else\n#define EMBM_TAB_ENTRIES GP_EXP_TAB"))
    text = open(src).read()
    sub = np.array([float(v) for v in re.findall(r"1\.[0-9]+", text[text.index("embm_sub_table[16] = {"):text.index("};", text.index("embm_sub_table[16] = {"))])])
    assert sub.shape == (16,) and np.max(np.abs(sub / np.exp2(np.arange(16) / 4096.0) - 1.0)) < 2.3e-16
    t12 = (t8[:, None] * sub[None, :]).reshape(-1)          # what every CTA builds: exp_tab[j] = T256[j >> 4] * T16[j & 15]
    assert abs(dm["EMBM_SCALE"] - 4096 / math.log(2)) < 1e-10 and abs(dm["EMBM_NEG_STEP"] + math.log(2) / 4096) < 1e-19
    y12 = _emulate(x, dm["EMBM_SCALE"], dm["EMBM_NEG_STEP"], 12, t12, [1.0, dm["EMBM_C1"], dm["EMBM_C2"]])
    e12 = float(np.max(np.abs(y12 / ref - 1)))
    print("max relative error of the device exp forms on [-700, 30]: 256/deg3 %.2e  64/deg4 %.2e  4096/deg2 %.2e" % (e8, e6, e12))
    # truncation + the argument reduction with one rounded ln2/N (|x| 1.1e-16)
    assert e8 < 1.2e-13 and e6 < 1.5e-13 and e12 < 1.2e-13
    small = np.abs(x) < 40.0
    assert float(np.max(np.abs(y8[small] / ref[small] - 1))) < 3e-14
    assert float(np.max(np.abs(y6[small] / ref[small] - 1))) < 5e-14
    assert float(np.max(np.abs(y12[small] / ref[small] - 1))) < 4e-14
