"""Regenerates profiles/r02_traffic.json -- DRAM bytes per launch of the dominant kernel of every bench.py workload --
from `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` logs of the benched build.
Run on the GPU box (tools/ncu_profiles.sh does), then commit the JSON: bench.py quotes it as roofline.traffic together
with the commit it was measured on (roofline.traffic_source).

    python tools/ncu_traffic.py <commit> <workload>,<kernel>,<trial_lik_per_launch>,<csv> ...
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
sys.path.insert(0, ROOT)
import bench  # noqa: E402

commit = bench.git_head() if sys.argv[1] in ("x", "auto") else sys.argv[1]
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for spec in sys.argv[2:]:
    workload, kernel, lik, path = spec.split(",", 3)
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    i_name, i_metric, i_unit, i_val = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    i_id = hdr.index("ID")
    per = {}
    for r in rows[1:]:
        if kernel.split("::")[-1] not in r[i_name]:
            continue
        v = float(r[i_val].replace(",", ""))
        u = r[i_unit].lower()
        if r[i_metric].startswith("dram__bytes"):
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        per.setdefault(r[i_id], {})[r[i_metric]] = v
    launches = [p for p in per.values() if "dram__bytes_read.sum" in p]
    if not launches:
        print(f"{spec}: no launch of {kernel} in {path}")
        continue
    # the largest launches are the ones bench.py's roofline is about (a warm-up or a migration-only launch is smaller)
    launches.sort(key=lambda p: -p.get("gpu__time_duration.sum", 0))
    top = launches[:max(1, len(launches) // 2)]
    dram = sum(p["dram__bytes_read.sum"] + p["dram__bytes_write.sum"] for p in top) / len(top)
    out.setdefault(workload, {})[kernel] = {"dram_bytes_per_launch": dram, "trial_lik_per_launch": float(lik), "launches_averaged": len(top),
                                            "commit": commit, "command": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none on bench.py --workload {workload}"}
    print(f"{workload} {kernel}: {dram / 1e6:.2f} MB per launch over {len(top)} launches")
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
