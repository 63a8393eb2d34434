#!/usr/bin/env python
"""profiles/traffic.json from the traffic_*.csv files of tools/gpu_profiles.sh: DRAM bytes (read + write) of one launch of
the hot kernel per workload / shard size / steps per launch, the key format bench.py looks up.

    python tools/make_traffic.py gpurun_out/<dir> > profiles/traffic.json"""
import csv
import glob
import json
import os
import re
import sys

out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes) from `ncu --metrics ... -c 1` of the hot kernel "
                   "inside bench.py's timed region (tools/gpu_profiles.sh); one launch each with ncu's cache control, so lines still "
                   "dirty in L2 at kernel end are not counted; `us` is that launch's gpu__time_duration under ncu"}
for path in sorted(glob.glob(os.path.join(sys.argv[1], "traffic_*.csv"))):
    m = re.match(r"traffic_(.+)_(\d+)_K(\d+)\.csv", os.path.basename(path))
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    if not m or len(rows) < 2:
        continue
    h = rows[0]
    vals = {r[h.index("Metric Name")]: (float(r[h.index("Metric Value")]), r[h.index("Metric Unit")]) for r in rows[1:]}

    def to_bytes(v, unit):
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    rd = to_bytes(*vals["dram__bytes_read.sum"])
    wr = to_bytes(*vals["dram__bytes_write.sum"])
    t, tu = vals["gpu__time_duration.sum"]
    key = "%s|%s|K=%s" % (m.group(1), m.group(2), m.group(3))
    out[key] = int(rd + wr)
    out[key + "|detail"] = {"read": int(rd), "write": int(wr), "us": t * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(tu, 1.0)}
print(json.dumps(out, indent=1))
