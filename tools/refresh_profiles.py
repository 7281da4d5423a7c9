#!/usr/bin/env python
"""Turn the scratch outputs of one GPU profiling run (gpurun_out/) into the tracked summaries under
profiles/ (development tool, run here after the gpurun call):

    python tools/refresh_profiles.py gpurun_out/prof_r01b.ncu-rep gpurun_out/launches_r01b.csv r01b

writes profiles/ncu_<kernel>_<tag>.txt (one per distinct kernel of the --set full capture),
profiles/ncu_traffic_<tag>.json (DRAM bytes per launch), profiles/launches_<tag>.csv plus a per-kernel
share table profiles/launch_shares_<tag>.txt, and profiles/sass_evidence_<tag>.txt.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import KEYS  # noqa: E402


def short(name):
    m = re.match(r"(?:void )?([A-Za-z0-9_]+)", name)
    return m.group(1) if m else name


def summaries(rep, tag, workload):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    seen, traffic = set(), {}
    # several launches of one kernel in the window (global_step_kernel: Kmm-only at set_globals, then the head;
    # psi2x_stats_kernel: the instantiation the device flag did not select returns at once): keep the longest
    best = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = short(d["Kernel Name"])
        t = float(d["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]
        if k not in best or t > best[k][0]:
            best[k] = (t, r)
    for k, (_, r) in best.items():
        d = dict(zip(hdr, r))
        lines = ["# ncu --set full --clock-control none summary, %s" % workload, "",
                 "kernel: %s  grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size"))]
        for key in KEYS:
            for h, u, v in zip(hdr, units, r):
                if h == key:
                    lines.append("  %-88s %14s %s" % (h, v, u))
        out = os.path.join(ROOT, "profiles", "ncu_%s_%s.txt" % (k, tag))
        open(out, "w").write("\n".join(lines) + "\n")

        def val(key):
            i = hdr.index(key)
            v, u = float(r[i]), units[i]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        traffic[k] = {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                      "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
                      "duration_ms": float(d["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[
                          units[hdr.index("gpu__time_duration.sum")]],
                      "source": "profiles/ncu_%s_%s.txt (ncu --set full --clock-control none, %s)" % (k, tag, workload)}
        print("wrote", out)
    return traffic


def launches(csv_path, tag):
    text = open(csv_path).read()
    body = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(body)))
    tot = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        e = tot.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += ns
    allns = sum(v[1] for v in tot.values())
    lines = ["# per-kernel share of the launch list profiles/launches_%s.csv" % tag,
             "# (ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3: 5 evaluations + set-up + probes;",
             "#  serialised, cold-cache per-launch times: compare SHARES with the live CUDA-event phases of bench.py)",
             "%-34s %8s %12s %7s" % ("kernel", "launches", "total ms", "share")]
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-34s %8d %12.3f %6.1f%%" % (k, n, ns / 1e6, 100.0 * ns / allns))
    open(os.path.join(ROOT, "profiles", "launch_shares_%s.txt" % tag), "w").write("\n".join(lines) + "\n")
    open(os.path.join(ROOT, "profiles", "launches_%s.csv" % tag), "w").write(body)
    print("\n".join(lines))


def sass(tag):
    lib = os.path.join(ROOT, "gparml_b200", "libgparml_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    want = ["prep_points_kernel", "psi1_mma_kernelILi10ELi1ELi2E", "psi2x_stats_kernelILi10ELb0E", "psi2_stats_kernelILi12E", "embed_psi2x_kernelILi10E", "embed_psi2m_kernelILi10E", "embed_psi2x_kernelILi12E",
            "embed_psi1_kernelILi10E", "global_step_kernel", "psi2_stats_f32_kernelILi10E", "embed_psi2_f32_kernelILi10E",
            "gsl_gemm_kernel", "scg_reduce_kernel", "scg_update_kernel", "init_scatter_kernel", "psi1_wide_kernelILi10E",
            "global_step_tail_kernel", "stats_allreduce_kernel", "gsl_panel_kernel", "gsl_grad_z_kernel"]
    mn = ["DFMA", "DADD", "DMUL", "DMMA", "DMMA.8x8x4", "DSETP", "FFMA", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "IMAD"]
    lines = ["# SASS evidence (cuobjdump -sass gparml_b200/libgparml_b200.so, sm_100a), %s" % tag,
             "# per kernel: instruction counts of the mnemonics that matter for this path",
             "#   DFMA/DADD/DMUL = FP64 pipe;  DMMA = FP64 tensor-core MMA (mma.sync m8n8k4 f64);  UBLKCP = cp.async.bulk (TMA unit,",
             "#   1-D bulk copy);  SYNCS = mbarrier;  LDGSTS = cp.async;  LDS/STS = shared memory;  MUFU = special function unit", ""]
    cur, counts = None, None
    funcs = collections.OrderedDict()
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        if cur:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if m:
                funcs[cur][m.group(1)] += 1
                funcs[cur]["_total"] += 1
    for f, c in funcs.items():
        if any(w in f for w in want):
            lines.append(f)
            lines.append("    total %d | " % c["_total"] + "  ".join("%s %d" % (k, c[k]) for k in mn if c[k]))
    open(os.path.join(ROOT, "profiles", "sass_evidence_%s.txt" % tag), "w").write("\n".join(lines) + "\n")
    print("wrote sass evidence,", len(lines), "lines")


if __name__ == "__main__":
    rep, lcsv, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    workload = "bench.py c3 N=1M, 1 GPU"
    tr = summaries(rep, tag, workload)
    for extra in sys.argv[4:]:                       # further reports of the same workload (e.g. the K6 kernels)
        tr.update(summaries(extra, tag, workload))
    for v in tr.values():
        v["n_local"] = 1000000
    json.dump({"c3": tr}, open(os.path.join(ROOT, "profiles", "ncu_traffic_%s.json" % tag), "w"), indent=1)
    launches(lcsv, tag)
    sass(tag)
