"""Print the interesting numbers of a bench.py JSON line: python tools/show_bench.py FILE"""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value %.3f evals/s  %.3f ms/step  e2e %s  driver %s  gpus %d" % (
    d["value"], d["ms_per_step"], d["e2e"] and "%.3f" % d["e2e"]["value"],
    d.get("through_driver") and "%.3f" % d["through_driver"]["value"], d["n_gpus"]))
print("phases", {k: round(v, 4) for k, v in d["phase_ms_median"].items()}, "F", d["F"], d.get("F_check"))
r = d["roofline"]
print("K2 frac %.3f exec %.3f | K5 frac %.3f exec %.3f | prep hbm %.3f | whole %.3f" % (
    r["frac"], r["executed_frac"], r.get("embed_grads", {}).get("frac", 0), r.get("embed_grads", {}).get("executed_frac", 0),
    r.get("prep_points_hbm", {}).get("frac", 0), r["whole_evaluation_frac"]))
k6 = r.get("scg_local_state_hbm")
if k6:
    print("K6 update_d %.0f GB/s (%.2f)  copy %.0f GB/s (%.2f)" % (k6["update_d"]["achieved"], k6["update_d"]["frac"],
                                                                 k6["update_grad_old"]["achieved"], k6["update_grad_old"]["frac"]))
if d.get("e2e") and d["e2e"].get("phase_ms_median"):
    print("e2e phases", {k: round(v, 3) for k, v in d["e2e"]["phase_ms_median"].items()}, d["e2e"].get("host_link_all_ranks_copying"))
print("clocks", d["clocks"], "allreduce", d.get("allreduce"), "launches", d["gpu_launches"])
for o in d.get("other_configs", []):
    print(" ", o["config"]["workload"].split(":")[0], "%.4f evals/s %.3f ms" % (o["value"], o["ms_per_step"]),
          "K2 %.3f/%.3f" % (o["psi2_stats"]["frac"], o["psi2_stats"]["executed_frac"]), "allreduce", o.get("allreduce"),
          "Fcheck", o.get("F_check") and o["F_check"]["ok"], {k: round(v, 3) for k, v in o["phase_ms_median"].items()},
          "clk", o["clocks"]["sm_mhz"], o["clocks"]["reasons"])
if d.get("cpu_baseline"):
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
