import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d.get("roofline") or {}; e=d.get("e2e") or {}
        print("%s it/s %.0f unfl %.0f ms %.4f launch_ms %.4f share %.2f frac %.3f e2e %.3e value %.3e"%(d["config"]["workload"][:3], d.get("iters_per_s",0), d.get("iters_per_s_unflushed",0), d["ms_per_step"], r.get("launch_ms",0), r.get("kernel_share_of_step",0), r.get("frac") or 0, e.get("value",0), d["value"]))
