#!/usr/bin/env python
"""Condensed view of bench.py JSON lines: show_bench.py FILE..."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as e:
        print("== %s: no JSON line (%r)" % (path, e))
        continue
    print("== %s" % path)
    if d.get("impl") == "reference":
        print("  reference arm: %.4g %s, %s" % (d["value"], d["unit"], d["cpu_baseline"]["sample"][:120]))
        continue
    r = d["roofline"]
    print("  n_gpus %d  %s  value %.4g  frac %.3f (layout %.3f)  region %.4f ms x %d  steps %d warmup %d  clocks %s" % (
        d["n_gpus"], d["config"]["workload"], d["value"], r["frac"], r["frac_this_layout"], d["region_ms"]["median"], d["repeats"],
        d["steps"], d["warmup"], d["clocks"]))
    o = d.get("other_mode")
    if o:
        print("  other mode %s: %.4g  frac %.3f (layout %.3f)" % (o["mode"], o["value"], o["roofline_frac"], o["roofline_frac_this_layout"]))
    print("  e2e: " + "  ".join("%s %.4g" % (k[4:] or "(headline: %d B/env)" % (v["d2h_bytes_per_step"] // d["config"]["total_envs"]), v["value"]) for k, v in d.items() if k.startswith("e2e") and v))
    if d.get("pcie"):
        print("  pcie d2h: %.1f GB/s per GPU, %.1f GB/s all" % (d["pcie"]["d2h_gbs_per_gpu"], d["pcie"]["d2h_gbs_all_gpus"]))
    for k, w in (d.get("workloads") or {}).items():
        if "failed" in w:
            print("  %-26s FAILED %s" % (k, w["failed"]))
            continue
        a, b = w["rollout"], w["step"]
        print("  %-26s rollout %.4g frac %.3f (%.3f) | step %.4g frac %.3f (%.3f) | L %.1f | allreduce %.3f ms" % (
            k, a["value"], a["frac"], a["frac_this_layout"], b["value"], b["frac"], b["frac_this_layout"],
            w["mean_episode_length"] or 0, w["stats_allreduce_ms"]))
    for k in ("cpu_baseline", "cpu_baseline_c"):
        if d.get(k):
            print("  %s: %s" % (k, {x: d[k].get(x) for x in ("value", "cores", "kind", "uncached_reset_ms")}))
    c1 = d.get("config1")
    if c1 and "failed" not in c1:
        print("  config1: cpu %s | ours %.4g env-steps/s" % (
            {k: round(v["env_steps_per_s"]) for k, v in c1["cpu"].items() if isinstance(v, dict)}, c1["ours"]["env_steps_per_s"]))
    elif c1:
        print("  config1 FAILED", c1)
