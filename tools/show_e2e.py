"""Print the e2e* entries of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.2f G  frac %.3f  n_gpus %d" % (d["value"] / 1e9, d["roofline"]["frac"], d["n_gpus"]))
for k, v in d.items():
    if k.startswith("e2e") and v:
        print("%-24s %8.1f M   d2h %d B/step" % (k, v["value"] / 1e6, v["d2h_bytes_per_step"]))
print(d["pcie"]["d2h_gbs_per_gpu"], d["pcie"]["d2h_gbs_all_gpus"])
