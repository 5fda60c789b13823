"""Reads a GGDMC_B200_ITEMTRACE file (stamps of the persistent sampler kernel's work items, last launch) and prints where the
time of an item goes: wait for its dependency, proposal, table build, trial loop, finish (arrive / MH decisions), plus the
launch's span and how busy the workers were.

    GGDMC_B200_ITEMTRACE=/tmp/t.bin python bench.py --workload c2 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline
    python tools/exp_itemtrace.py /tmp/t.bin
"""
import os
import sys
import numpy as np

if sys.argv[1] == "run":  # run <workload> <iterations per launch> <file> [subjects]: writes the trace of one launch of that many iterations
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["GGDMC_B200_ITEMTRACE"] = sys.argv[4]
    os.environ["GGDMC_B200_BATCH"] = sys.argv[3]
    import bench
    from ggdmc_b200 import engine as E, workloads as W
    model_k, S, ntr, _ = bench.WORKLOADS[sys.argv[2]]
    if len(sys.argv) > 5:
        S = int(sys.argv[5])
    w = W.hierarchical(sys.argv[2], model_k, S, ntr, n_replicate=1, subject_begin=0, subject_end=S)
    tun = W.tuning_for(w, nmc=2, thin=1 << 30, seeds=[9032], pop_migration_prob=0.05, sub_migration_prob=0.05, subject_begin=0, n_subject_total=S)
    eng = E.Engine(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, w.subj_start)
    n = int(sys.argv[3])
    for _ in range(3):
        ms = eng.iterate(n)
    print(f"{sys.argv[2]} ({S} subjects): {n} iterations in one launch: {ms:.3f} ms = {n / ms * 1e3:.1f} iterations/s")
    eng.close()
    sys.argv[1] = sys.argv[4]

raw = np.fromfile(sys.argv[1], dtype=np.uint64)
cap, npop, nslot, nsplit, n_phi, grid, threads, per_iter = (int(x) for x in raw[:8])
rec = raw[8:8 + 8 * cap].reshape(-1, 8)
rec = rec[rec[:, 0] > 0]
kind = (rec[:, 7] & np.uint64(0xff)).astype(int)
it = (rec[:, 7] >> np.uint64(8)).astype(np.int64)
it -= it.min()
warp_slot = (rec[:, 6] >> np.uint64(16)) & np.uint64(0xffff)
rec[:, 6] &= np.uint64(0xffff)
t0 = rec[:, 0].min()
names = {0: "SUBJECT h0", 1: "SUBJECT h1", 2: "PHI h0", 3: "PHI h1", 4: "CLOSE"}
print(f"{len(rec)} items traced (cap {cap}); {npop} populations x {nslot} slots x {nsplit} chunks; {n_phi} phi items per half; "
      f"grid {grid} x {threads} threads = {grid * threads // 32} workers")
end = rec[:, 5].astype(np.int64)
span = (end.max() - int(t0)) / 1e3
print(f"span of the traced items: {span:.1f} us")


def us(a):
    a = np.asarray(a, dtype=np.float64) / 1e3
    return f"mean {a.mean():7.2f}  p50 {np.median(a):7.2f}  p95 {np.percentile(a, 95):7.2f}  max {a.max():7.2f}"


busy = 0.0
for k in sorted(names):
    r = rec[kind == k]
    if not len(r):
        continue
    r = r.astype(np.int64)
    print(f"--- {names[k]}: {len(r)} items, first taken at {(r[:, 0].min() - int(t0)) / 1e3:.1f} us, last finished at {(r[:, 5].max() - int(t0)) / 1e3:.1f} us")
    if k < 2:
        print(f"    wait for dependency : {us(r[:, 1] - r[:, 0])}")
        ok = r[:, 2] > 0
        rr = r[ok]
        if len(rr):
            print(f"    proposal            : {us(rr[:, 2] - rr[:, 1])}")
            print(f"    table build         : {us(rr[:, 3] - rr[:, 2])}")
            print(f"    trial loop          : {us(rr[:, 4] - rr[:, 3])}")
            print(f"    finish              : {us(rr[:, 5] - rr[:, 4])}")
            busy += float((rr[:, 5] - rr[:, 1]).sum())
    elif k < 4:
        print(f"    proposal + sums     : {us(r[:, 4] - r[:, 1])}")
        print(f"    finish              : {us(r[:, 5] - r[:, 4])}")
        busy += float((r[:, 5] - r[:, 1]).sum())
    else:
        print(f"    MH decisions        : {us(r[:, 5] - r[:, 1])}")
        busy += float((r[:, 5] - r[:, 1]).sum())
workers = grid * threads // 32
print(f"worker time in items (waits excluded): {busy / 1e3:.0f} us = {100 * busy / 1e3 / (span * workers):.1f} % of {workers} workers x span")
for i in np.unique(it)[1:3]:
    m = it == i
    for k in sorted(names):
        mk = m & (kind == k)
        if mk.any():
            r = rec[mk].astype(np.int64)
            print(f"   iteration +{i} {names[k]:11s}: {mk.sum():6d} items, started {(r[:, 1].min() - int(t0)) / 1e3:8.1f} .. {(r[:, 1].max() - int(t0)) / 1e3:8.1f} us, finished "
                  f"{(r[:, 5].min() - int(t0)) / 1e3:8.1f} .. {(r[:, 5].max() - int(t0)) / 1e3:8.1f} us")
