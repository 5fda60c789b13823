import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("it/s %.0f unfl %.0f ms %.4f like_ms %.4f share %.2f frac %.3f"%(d["iters_per_s"], d["iters_per_s_unflushed"], d["ms_per_step"], d["roofline"]["launch_ms"], d["roofline"]["kernel_share_of_step"], d["roofline"]["frac"]))
