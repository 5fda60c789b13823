"""Diagnostics: per-block timeline of one likelihood launch (GGDMC_B200_BLOCKTRACE).  Usage: python tools/exp_blocktrace.py S"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
path = f"/tmp/bt_{S}.bin"
os.environ["GGDMC_B200_BLOCKTRACE"] = path
from ggdmc_b200 import _lib as B, engine as E, workloads as W
w = W.hierarchical("c4", 6, S, 768, n_replicate=1, subject_begin=0, subject_end=S)
tun = W.tuning_for(w, nmc=2, thin=1 << 30, seeds=[9032], schedule=B.SCHEDULE_PARALLEL, subject_begin=0, n_subject_total=S, device=0)
eng = E.Engine(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, w.subj_start)
ms, n = eng.time_likelihood(10)
print(f"S={S}: {ms*1e3:.1f} us per launch (events, 10 launches back to back), {n} trial-likelihoods")
h = np.fromfile(path, dtype=np.uint64)
t0, t1 = int(h[0]), int(h[1])
b = h[2:].reshape(-1, 5).astype(np.int64)
st, en, sm = b[:, 0] - t0, b[:, 1] - t0, b[:, 2]
print(f"blocks {len(b)}; kernel bracket (stamp kernels) {1e-3*(t1-t0):.1f} us; first block start {st.min()*1e-3:.1f} us, last block end {en.max()*1e-3:.1f} us")
dur = (en - st) * 1e-3
pro, loop, epi = (b[:, 3] - b[:, 0]) * 1e-3, (b[:, 4] - b[:, 3]) * 1e-3, (b[:, 1] - b[:, 4]) * 1e-3
print(f"  per block (thread 0's view): table build {pro.mean():.2f} us, trial loop {loop.mean():.2f} us, reduction + exit {epi.mean():.2f} us")
order = np.argsort(st)
nb = len(b)
for lo, hi in [(0, 0.1), (0.1, 0.35), (0.35, 0.7), (0.7, 0.9), (0.9, 1.0)]:
    idx = order[int(lo * nb):int(hi * nb)]
    print(f"  blocks by start order {lo:.2f}-{hi:.2f}: start {st[idx].min()*1e-3:7.1f}..{st[idx].max()*1e-3:7.1f} us, duration mean {dur[idx].mean():6.1f} min {dur[idx].min():6.1f} max {dur[idx].max():6.1f} us")
# active blocks over time
ts = np.linspace(0, en.max(), 41)
act = [(int(((st <= t) & (en > t)).sum())) for t in ts]
print("  active blocks at 40 evenly spaced times:", act)
per_sm_end = np.array([en[sm == s].max() for s in np.unique(sm)]) * 1e-3
print(f"  per-SM finish time: min {per_sm_end.min():.1f} median {np.median(per_sm_end):.1f} max {per_sm_end.max():.1f} us; SMs used {len(per_sm_end)}")
