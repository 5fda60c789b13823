#!/usr/bin/env python
"""Benchmark of the ggdmc hot path on B200: DE-MCMC iterations of the hierarchical LBA fit.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one DE-MCMC iteration (src/de.cpp:281-381: one phi step + one step of every chain of
every subject) of BASELINE config 4 -- 1024 synthetic subjects x 768 trials, 78 chains, the README
B x v model -- with subjects sharded over the N GPUs (strong scaling: total work fixed).
`--workload c1 | c2 | c3 | c5` runs the other BASELINE configs with the same JSON line (c1: the README single-subject
fit, a step = one iteration of run_chains, src/de.cpp:208-240, three replicates like the README's ncore = 3; c3: the
likelihood-only sweep, a step = one sum-log-likelihood pass over --trials trials x 15 chains); `--sustain-s T` repeats
the timed loop for at least T seconds with clocks and power sampled every 20 ms.

  value     trial-likelihoods/s with data, state and sample storage resident in HBM; the number of
            trial-likelihoods is counted on the device (migration sweeps evaluate fewer chains); L2 flushed
            before every timed iteration, CUDA events, max over ranks
  e2e       the same metric through the reference-facing call `ggdmc_b200_run` (the C-ABI twin of
            .Call("_ggdmc_run")) with HOST buffers: upload of data + start state, K iterations, every stored
            sample copied back to the host arrays (streamed behind the sampler), all inside the timed region
  roofline  the dominant kernel against the FP64 FMA peak measured on this GPU by a DFMA microbenchmark
            (MEASURED_PEAKS.json has no FP64 entry); achieved = 513 algorithmic flop per 2-accumulator
            trial-likelihood x trial-likelihoods per launch / mean CUDA-event duration of the
            launches of a second pass over K more iterations.  The dominant kernel of the default schedule is
            the persistent sampler kernel gg::k_sampler (one launch = one whole iteration here: proposals, row
            tables, trial loops, MH decisions, phi step, storage); with GGDMC_B200_NO_PERSIST=1, and for c3, it is
            the likelihood kernel gg::k_like, every launch timed alone
  cpu_baseline  de_class::run_hchains of the reference's OWN object code (src/de.o behind an R-API shim,
            oracle/_ref; kind "reference") on one host core on a bounded sample of the same workload, with the
            -O2 C restatement (the oracle, kind "port") beside it; the port alone where oracle/_ref is absent

`--impl reference` times that reference CPU path with one replicate per host core, like the reference's
`ncore` forked replicates (R/sampling.R:26-55), on a bounded sample; rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

F_TRIAL = {2: 513.0, 3: 755.0, 4: 997.0}  # algorithmic flop per trial-likelihood: 242 * n_acc + 29
METRIC = "trial-likelihoods/s, DE-MCMC iterations of a hierarchical LBA fit"
UNIT = "trial-likelihoods/s"
W_NPAR = {6: 13, 5: 17, 2: 8}  # free parameters of the fixture models

WORKLOADS = {
    # name: (model fixture, subjects, trials per subject, description)
    "c1": (6, 1, 768, "C1: README single-subject LBA B x v model (13 par, 24 cells, 2 acc), 768 trials, 39 chains, 3 replicates "
                      "(ncore = 3), sub migration 0.06"),
    "c3": (0, 1, 0, "C3: likelihood-only throughput, 2-accumulator LBA (5 par, 2 cells), --trials trials x 15 chains"),
    "c4": (6, 1024, 768, "C4: hierarchical LBA B x v model (13 par, 24 cells, 2 acc), 1024 subjects x 768 trials, 78 chains, "
                         "pop+sub migration 0.05"),
    "c2": (6, 32, 768, "C2: README hierarchical recovery study, 32 subjects x 768 trials, 78 chains, pop+sub migration 0.05"),
    "c5": (5, 256, 2048, "C5: 4-accumulator LBA, 96 cells, 17 par, 256 subjects x 2048 trials, 102 chains, pop+sub migration 0.05"),
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_hier_sample(model_k: int, n_subject: int, n_trial: int, n_iter: int, seed: int, n_warm: int = 0):
    """Run the CPU oracle's run_hchains on `n_subject` synthetic subjects; returns (seconds, trial-likelihoods)."""
    from ggdmc_b200 import synth
    from ggdmc_b200.workloads import load_model
    from oracle import binding as ob

    spec = load_model(model_k)
    ct = spec.ct
    D, C = ct.npar, 6 * ct.npar
    om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar)

    def opr(p):
        return ob.OPrior(p.p0, p.p1, p.lower, p.upper, p.dist, p.log_p)

    opp, ohp = opr(spec.p_prior), opr(spec.h_prior)
    rng = np.random.default_rng([20260101, seed])
    center = np.concatenate([spec.pop_mean, spec.pop_scale])
    phi0 = np.abs(center[None, :] * (1.0 + 0.05 * rng.standard_normal((C, 2 * D))))
    datas, pops = [], []
    for s in range(n_subject):
        th = synth.rtnorm(spec.pop_mean, spec.pop_scale, 0.0, rng)
        tr = synth.simulate_subject(ct, spec.node_1_index, th, n_trial, rng)
        od = ob.OData(tr.rt, tr.cell)
        x0 = np.abs(th[None, :] * (1.0 + 0.05 * rng.standard_normal((C, D))))
        lp = np.array([ob.sumlogprior(opp, x0[c], phi0[c, :D], phi0[c, D:]) for c in range(C)])
        ll = np.array([ob.sumloglike(om, od, x0[c]) for c in range(C)])
        datas.append(od)
        pops.append(ob.OPop(x0, lp, ll, 2, 1 << 30))
    phi = ob.OPop(phi0, np.array([ob.sumlogprior(ohp, phi0[c]) for c in range(C)]), np.zeros(C), 2, 1 << 30)
    de = ob.make_de(2 * D, C, pop_migration_prob=0.05, sub_migration_prob=0.05, jacobi=0)
    r = ob.make_rng(seed=seed)
    if n_warm:
        ob.run_hier(de, phi, pops, opp, ohp, om, datas, r, n_warm)
    t0 = time.perf_counter()
    ob.run_hier(de, phi, pops, opp, ohp, om, datas, r, n_iter)
    dt = time.perf_counter() - t0
    # crossover sweeps evaluate every chain; the 5 % migration sweeps fewer -- count the nominal
    # crossover figure scaled by the expected fraction (0.95 + 0.05 * ~0.5)
    n_lik = n_iter * n_subject * C * n_trial * (0.95 + 0.05 * 0.5)
    return dt, n_lik


def refobj_hier_sample(model_k: int, n_subject: int, n_trial: int, n_iter: int, seed: int):
    """The reference's OWN machine code -- de_class::run_hchains of src/de.o (the package author's build,
    -O0) through oracle/_ref -- on `n_subject` synthetic subjects; returns (seconds, trial-likelihoods)."""
    from ggdmc_b200 import synth
    from ggdmc_b200.workloads import load_model
    from oracle import binding as ob

    spec = load_model(model_k)
    ct = spec.ct
    D, C = ct.npar, 6 * ct.npar
    om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar)

    def opr(p):
        return ob.OPrior(p.p0, p.p1, p.lower, p.upper, p.dist, p.log_p)

    opp, ohp = opr(spec.p_prior), opr(spec.h_prior)
    rng = np.random.default_rng([20260101, seed])
    center = np.concatenate([spec.pop_mean, spec.pop_scale])
    phi0 = np.abs(center[None, :] * (1.0 + 0.05 * rng.standard_normal((C, 2 * D))))
    datas, starts = [], []
    for s in range(n_subject):
        th = synth.rtnorm(spec.pop_mean, spec.pop_scale, 0.0, rng)
        tr = synth.simulate_subject(ct, spec.node_1_index, th, n_trial, rng)
        od = ob.OData(tr.rt, tr.cell)
        x0 = np.abs(th[None, :] * (1.0 + 0.05 * rng.standard_normal((C, D))))
        lp = np.array([ob.sumlogprior(opp, x0[c], phi0[c, :D], phi0[c, D:]) for c in range(C)])
        ll = np.array([ob.sumloglike(om, od, x0[c]) for c in range(C)])
        datas.append(od)
        starts.append((x0, lp, ll))
    phi_s = (phi0, np.array([ob.sumlogprior(ohp, phi0[c]) for c in range(C)]), np.zeros(C))
    ob.ref2_prime()
    ob.ref_lib().ref_set_uniform_stream(None, 0)  # Rf_runif falls back to the shim's own generator
    t0 = time.perf_counter()
    ob.ref2_run_hchains(2 * D, om, datas, opp, ohp, phi_s, starts, n_iter + 1, 1, pop_migration_prob=0.05, sub_migration_prob=0.05)
    dt = time.perf_counter() - t0
    n_lik = n_iter * n_subject * C * n_trial * (0.95 + 0.05 * 0.5)
    return dt, n_lik


def _ref_worker(args):
    model_k, n_subject, n_trial, steps, warm, seed, use_obj = args
    if use_obj:
        if warm:
            refobj_hier_sample(model_k, n_subject, n_trial, warm, seed + 7)
        return refobj_hier_sample(model_k, n_subject, n_trial, steps, seed)
    return oracle_hier_sample(model_k, n_subject, n_trial, steps, seed, n_warm=warm)


def run_reference(args, rank, world):
    """The reference's CPU path on the host cores, bounded sample.  With oracle/_ref present this is the
    reference's own object code (de_class::run_hchains of src/de.o); otherwise the oracle port.  One replicate
    process per host core, like the reference's `ncore` forked replicates (R/sampling.R:26-55)."""
    if rank != 0:
        return
    import multiprocessing as mp

    model_k, S, ntr, desc = WORKLOADS[args.workload]
    cores = max(1, min(os.cpu_count() or 1, 64))
    n_sub = 4  # bounded sample: 4 of the subjects per replicate process
    from oracle import binding as ob
    ob.build()
    use_obj = ob.ref_lib() is not None
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_ref_worker, [(model_k, n_sub, ntr, args.steps, args.warmup, 1000 + i, use_obj) for i in range(cores)])
    wall = time.perf_counter() - t0
    tmax = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    value = total / tmax
    nchain = 6 * W_NPAR[model_k]
    kind = "reference" if use_obj else "port"
    what = ("the reference's own object code: de_class::run_hchains of src/de.o (package author's build, -O0) behind an R-API shim"
            if use_obj else "oracle C port at -O2")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tmax / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "sample": f"{n_sub} subjects x {ntr} trials x {nchain} chains per replicate process"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{cores} replicate processes (one per host core, like the reference's ncore forks), each "
                                   f"{args.steps} DE-MCMC iterations over {n_sub} subjects of the workload; {what}; wall {wall:.1f} s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def git_head() -> str:
    """Identity of the benched build: the commit where git is available, and always the hash of the kernel sources (the
    GPU box gets a snapshot without .git)."""
    import hashlib
    h = hashlib.sha1()
    src = os.path.join(ROOT, "ggdmc_b200", "csrc")
    for f in sorted(os.listdir(src)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(src, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "ggdmc_b200.h"), "rb").read())
    ident = "src-" + h.hexdigest()[:12]
    try:
        c = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=5).stdout.strip()
        if c:
            ident = c + "/" + ident
    except Exception:
        pass
    return ident


def committed_traffic(workload: str, kernel: str, lik_per_launch: float):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/r02_traffic.json,
    regenerated by tools/ncu_traffic.sh whenever a kernel changes) -> (bytes or None, where it came from)."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        e = tj[workload][kernel]
        if abs(e["trial_lik_per_launch"] - lik_per_launch) < 0.03 * lik_per_launch:
            same = e.get("commit", "?").split("/")[-1] == git_head().split("/")[-1]
            return e["dram_bytes_per_launch"], (f"profiles/r02_traffic.json: ncu capture of build {e.get('commit', '?')} ({e.get('command', '')}); this run is build "
                                                f"{git_head()} ({'the same kernel sources' if same else 'DIFFERENT kernel sources: regenerate with tools/ncu_profiles.sh'})")
    except Exception:
        pass
    return None, "no committed ncu capture matches this workload / kernel / launch size"


def cpu_baseline_line(model_k: int, ntr: int, nchain: int):
    """The reference's CPU path on ONE host core on a bounded sample of the hierarchical workload: de_class::run_hchains of
    the reference's own object code (kind "reference") with the -O2 oracle port beside it."""
    from oracle import binding as ob
    ob.build()
    dt1, n1 = oracle_hier_sample(model_k, 2, ntr, 1, 7)
    iters = int(max(1, min(50, 12.0 / max(dt1 * 4, 1e-3))))  # aim at ~12 s of CPU work on 8 subjects
    dtc, nc = oracle_hier_sample(model_k, 8, ntr, iters, 8)
    port = {"value": nc / dtc, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{iters} DE-MCMC iterations (run_hchains restatement, reference chain order) over 8 subjects x {ntr} trials x "
                      f"{nchain} chains of the same synthetic population; {dtc:.1f} s on one host core, gcc -O2"}
    if ob.ref_lib() is None:
        return port
    dtr, nr = refobj_hier_sample(model_k, 4, ntr, 2, 9)
    it_r = int(max(2, min(40, 10.0 / max(dtr / 2, 1e-3))))
    dtr, nr = refobj_hier_sample(model_k, 4, ntr, it_r, 10)
    return {"value": nr / dtr, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"{it_r} DE-MCMC iterations of de_class::run_hchains from the reference's own src/de.o (package author's "
                      f"build, -O0; R-API shim: Cody pnorm, injected runif) over 4 subjects x {ntr} trials x {nchain} chains; "
                      f"{dtr:.1f} s on one host core",
            "port_value": port["value"], "port_cores": 1, "port_note": "the same path as the -O2 C restatement (oracle), one core", "port": port}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--schedule", default="parallel", choices=["parallel", "reference", "simultaneous"])
    ap.add_argument("--subjects", type=int, default=0, help="experiments only: override the workload's subject count")
    ap.add_argument("--trials", type=int, default=1_000_000, help="c3 only: trials of the likelihood-only sweep point")
    ap.add_argument("--migration", type=float, default=0.05, help="experiments only: pop and sub migration probability")
    ap.add_argument("--sustain-s", type=float, default=0.0, help="repeat the timed loop for at least this many seconds (clocks under sustained load)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.workload in ("c1", "c3"):
            raise SystemExit("--impl reference times the hierarchical path (c2, c4, c5); tools/exp_c1.py has the single-subject reference timing")
        run_reference(args, rank, world)
        return

    from ggdmc_b200 import _lib as B
    from ggdmc_b200 import engine as E
    from ggdmc_b200 import workloads as W

    B.build()
    if E.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: ggdmc_b200 has no CPU fallback")
    if args.workload in ("c1", "c3") and world > 1:
        raise SystemExit("c1 and c3 are single-GPU configurations (a single-subject fit stays on one GPU)")

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
        uid = [E.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        E.comm_init(world, rank, uid[0], local_rank)
    else:
        import ctypes
        # make device `local_rank` current for this process without torch
        cudart = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
        if cudart is not None and local_rank:
            cudart.cudaSetDevice(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    schedule = {"parallel": B.SCHEDULE_PARALLEL, "reference": B.SCHEDULE_REFERENCE, "simultaneous": B.SCHEDULE_SIMULTANEOUS}[args.schedule]
    fp64_peak = E.measure_fp64_tflops(local_rank) if rank == 0 else 0.0
    K, Wm = args.steps, args.warmup
    seeds = [9032]

    if args.workload == "c3":
        bench_c3(args, E, W, fp64_peak)
        return

    # ---- the resident engine of the workload ---------------------------------------------------
    model_k, S, ntr, desc = WORKLOADS[args.workload]
    if args.workload == "c1":
        from ggdmc_b200 import synth
        spec = W.load_model(6)
        ct = spec.ct
        D, nchain, R = ct.npar, 3 * ct.npar, 3
        seeds = [9032, 9033, 9034]
        rng = np.random.default_rng(20260101)
        theta_true = synth.rtnorm(spec.pop_mean, spec.pop_scale, 0.0, rng)
        tr = synth.simulate_subject(ct, spec.node_1_index, theta_true, ntr, rng)
        x0 = np.abs(theta_true[None, None, :] * (1.0 + 0.05 * rng.standard_normal((R, nchain, D))))
        ll0 = E.sumloglike(ct, [tr], x0.reshape(1, R * nchain, D)).reshape(R, nchain)
        lp0 = E.sumlogprior(spec.sub_prior, x0.reshape(R * nchain, D)).reshape(R, nchain)
        start = E.PopState(x0, lp0, ll0)
        s0, s1 = 0, 1

        def make_tuning(nmc, thin):
            return E.Tuning(nmc=nmc, nchain=nchain, thin=thin, nparameter=D, sub_migration_prob=0.06, seeds=seeds, schedule=schedule, device=local_rank)

        def make_engine():
            return E.Engine(ct, [tr], spec.sub_prior, None, make_tuning(2, 1 << 30), None, [start])
        n_acc = ct.n_acc
    else:
        if args.subjects > 0:
            S, desc = args.subjects, desc + f" [EXPERIMENT: {args.subjects} subjects]"
        s0, s1 = W.shard_bounds(S, rank, world)
        w = W.hierarchical(args.workload, model_k, S, ntr, n_replicate=1, subject_begin=s0, subject_end=s1)
        n_acc, nchain = w.spec.ct.n_acc, w.nchain

        def make_engine():
            tun = W.tuning_for(w, nmc=2, thin=1 << 30, seeds=seeds, schedule=schedule, pop_migration_prob=args.migration,
                               sub_migration_prob=args.migration, subject_begin=s0, n_subject_total=S, device=local_rank)
            return E.Engine(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, w.subj_start)
    eng = make_engine()
    persistent = eng.persistent

    eng.iterate(Wm)
    eng.counters()
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall = time.perf_counter()
    ms = eng.iterate_flushed(K, 256 << 20)
    n_timed = K
    while args.sustain_s > 0 and time.perf_counter() - t_wall < args.sustain_s:  # sustained load: keep going, same protocol
        ms += eng.iterate_flushed(K, 256 << 20)
        n_timed += K
    clocks = sampler.stop()
    barrier()
    n_lik_local, _, _ = eng.counters()
    launches = eng.launch_count - launches0
    ms_max = allmax(ms)
    n_lik = allsum(float(n_lik_local))
    value = n_lik / (ms_max * 1e-3)
    # the same K iterations back to back without the flushes (how a fit actually runs: L2 stays warm, the ranks stay in
    # lock-step through the exchange, and the persistent kernel runs many iterations per launch) -- reported beside
    # `value`, never instead of it
    barrier()
    ms_warm = allmax(eng.iterate(K))
    eng.counters()
    # second pass over K more iterations with the dominant kernel's launches bracketed by CUDA events on the engine's stream
    eng.profile(True)
    ms_prof = eng.iterate_flushed(K, 256 << 20)
    n_lik_prof, like_ms, like_launches = eng.counters()
    eng.profile(False)

    # ---- roofline of the dominant kernel (rank 0's launches) -----------------------------------
    roofline = None
    if rank == 0 and like_launches > 0:
        kernel = "gg::k_sampler" if persistent else "gg::k_like"
        per_launch_s = like_ms * 1e-3 / like_launches
        lik_per_launch = n_lik_prof / like_launches
        achieved = F_TRIAL[n_acc] * lik_per_launch / per_launch_s / 1e12
        bytes_per_launch = 10.0 * lik_per_launch
        traffic, traffic_source = committed_traffic(args.workload if world == 1 and args.subjects == 0 and args.schedule == "parallel" else "-", kernel, lik_per_launch)
        roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak if fp64_peak > 0 else None, "traffic": traffic, "traffic_source": traffic_source,
                    "kernel": kernel, "peak_source": "DFMA microbenchmark on this GPU (ggdmc_b200_measure_fp64_tflops), burst",
                    "flop_per_trial_lik": F_TRIAL[n_acc], "trial_lik_per_launch": lik_per_launch,
                    "launch_ms": per_launch_s * 1e3, "kernel_share_of_step": like_ms / ms_prof,
                    "hbm_side": {"algorithmic_GBps": bytes_per_launch / per_launch_s / 1e9, "bytes_per_trial_lik": 10}}

    # ---- end to end through the reference-facing call with host buffers -----------------------
    # (the resident engine goes first: its device memory returns to the process's pool, as it does between the stages of a fit)
    eng.close()
    e2e = None
    if not args.no_e2e:
        thin = max(d for d in range(1, 9) if K % d == 0)  # the reference's README fits use thin = 8 (README.md:181-196)
        nmc = K // thin + 1
        if args.workload == "c1":
            E.run_subject(ct, tr, spec.sub_prior, make_tuning(2, 1), start)  # the call's one-off costs (module load) are not the fit's
            t0 = time.perf_counter()
            out1 = E.run_subject(ct, tr, spec.sub_prior, make_tuning(nmc, thin), start)
            dt_max = time.perf_counter() - t0
            h2d = tr.rt.nbytes + tr.cell.nbytes + x0.nbytes + lp0.nbytes + ll0.nbytes
            d2h = out1.theta.nbytes + out1.lp.nbytes + out1.ll.nbytes
            assert np.all(np.isfinite(out1.theta))
            call = "ggdmc_b200_run_subject (C-ABI twin of .Call('_ggdmc_run_subject'))"
        else:
            tun2 = W.tuning_for(w, nmc=nmc, thin=thin, seeds=seeds, schedule=schedule, subject_begin=s0, n_subject_total=S, device=local_rank)
            # caller-owned result arrays, allocated (and touched) before the timed region like any reused buffer
            outs = E.alloc_hier_outputs(len(w.trials), 1, nmc, w.nchain, w.spec.ct.npar, touch=True)
            barrier()
            t0 = time.perf_counter()
            phi_out, subj_out = E.run_hier(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun2, w.phi_start, w.subj_start, out=outs)
            dt_max = allmax(time.perf_counter() - t0)
            h2d = sum(t.rt.nbytes + t.cell.nbytes for t in w.trials) + sum(s.theta.nbytes + s.lp.nbytes + s.ll.nbytes for s in w.subj_start)
            h2d += w.phi_start.theta.nbytes + w.phi_start.lp.nbytes + w.phi_start.ll.nbytes
            d2h = sum(o.theta.nbytes + o.lp.nbytes + o.ll.nbytes for o in subj_out) + phi_out.theta.nbytes + phi_out.lp.nbytes + phi_out.ll.nbytes
            assert np.all(np.isfinite(phi_out.theta))
            call = "ggdmc_b200_run (C-ABI twin of .Call('_ggdmc_run'))"
        e2e = {"value": (n_lik / n_timed * K) / dt_max, "unit": UNIT, "h2d_bytes_per_step": allsum(h2d) / K, "d2h_bytes_per_step": allsum(d2h) / K,
               "ms_per_step": 1e3 * dt_max / K, "thin": thin, "nmc": nmc,
               "call": f"{call} with pageable host buffers: upload of trials + start "
                       f"state, {K} iterations storing every {thin}th (nmc = {nmc}), every stored sample copied to the host arrays (streamed one slot behind the sampler); host wall "
                       "clock around the call, max over ranks; trial-likelihoods per iteration from the resident phase's device counter"}

    # ---- CPU baseline: the reference's path on one host core, bounded sample --------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_line(6 if args.workload == "c1" else model_k, ntr, 78 if args.workload == "c1" else nchain)
        if args.workload == "c1":
            cpu["note"] = "timed on the hierarchical form of the same model (run_hchains); tools/exp_c1.py times de.o's run_chains on this very fit"

    if rank == 0:
        timing = ("CUDA events on the engine stream around each iteration, summed; max over ranks. One iteration = ONE launch of the persistent "
                  "sampler kernel (every warp a worker on a device-side queue of proposal / phi / close items; gg_sampler.cuh). roofline pass: "
                  "the same launches bracketed one by one") if persistent else \
                 ("CUDA events on the engine stream around each iteration (one CUDA-graph launch: phi sweep on a side stream, two subject groups "
                  "on their own streams, all joined before the closing event), summed; max over ranks. roofline pass: same iterations as plain "
                  "launches on ONE stream so that every k_like launch is timed alone")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_max / n_timed,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "schedule": args.schedule, "subjects_per_gpu": s1 - s0, "nchain": nchain,
                       "trials_per_subject": ntr, "l2": "flushed: 256 MiB memset before every timed iteration, outside the event brackets (at N > 1 followed by a peer-memory barrier, also outside, so the memsets' skew is not booked as exchange wait)",
                       "timing": timing, "seeds": seeds, "persistent_kernel": bool(persistent), "commit": git_head()},
            "iters_per_s": n_timed / (ms_max * 1e-3),
            "iters_per_s_unflushed": K / (ms_warm * 1e-3),
            "trial_lik_per_iter": n_lik / n_timed,
            "timed_iterations": n_timed,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
        }
        if args.sustain_s > 0:
            line["sustained_s"] = time.perf_counter() - t_wall
        print(json.dumps(line), flush=True)
    if dist is not None:
        E.comm_finalize()
        dist.destroy_process_group()


def bench_c3(args, E, W, fp64_peak):
    """BASELINE config 3: one sum-log-likelihood pass over --trials trials x 15 chains of the 5-parameter 2-accumulator
    model (gg::k_like alone); e2e = ggdmc_b200_sumloglike with host buffers (upload, pass, download)."""
    from ggdmc_b200 import synth
    from ggdmc_b200.model import Trials
    ct, node_1, p_vector, prior = W.sweep_model()
    rng = np.random.default_rng(20260103)
    base = synth.simulate_subject(ct, node_1, p_vector, 100_000, rng)
    nchain, n = 15, int(args.trials)
    tr = Trials(base.rt[:: 100_000 // n][:n].copy(), base.cell[:: 100_000 // n][:n].copy()) if n <= 100_000 else \
        Trials(np.tile(base.rt, n // 100_000), np.tile(base.cell, n // 100_000))
    order = np.argsort(tr.cell, kind="stable")
    tr = Trials(tr.rt[order], tr.cell[order])
    theta = p_vector * (1.0 + 0.05 * rng.uniform(-1, 1, size=(1, nchain, 5)))
    ll = E.sumloglike(ct, [tr], theta[0][None])[0]
    lp = E.sumlogprior(prior, theta[0])
    tun = E.Tuning(nmc=2, nchain=nchain, thin=1 << 30, nparameter=5, seeds=[1])
    eng = E.Engine(ct, [tr], prior, None, tun, None, [E.PopState(theta, lp[None], ll[None])])
    K, Wm = args.steps, args.warmup
    eng.time_likelihood(Wm)
    sampler = ClockSampler(0)
    sampler.start()
    ms, nlik = eng.time_likelihood(K)  # mean ms per launch; inputs (8 + 2 B per trial) larger than L2 from 1.3e7 trials on
    clocks = sampler.stop()
    eng.close()
    e2e = None
    if not args.no_e2e:
        E.sumloglike(ct, [tr], theta[0][None])
        t0 = time.perf_counter()
        for _ in range(K):
            E.sumloglike(ct, [tr], theta[0][None])
        dt = (time.perf_counter() - t0) / K
        e2e = {"value": nlik / dt, "unit": UNIT, "h2d_bytes_per_step": tr.rt.nbytes + tr.cell.nbytes + theta.nbytes, "d2h_bytes_per_step": 8 * nchain,
               "ms_per_step": 1e3 * dt, "call": "ggdmc_b200_sumloglike with pageable host buffers: upload of the trials, one pass, download of the sums"}
    value = nlik / (ms * 1e-3)
    achieved = F_TRIAL[2] * value / 1e12
    traffic, traffic_source = committed_traffic("c3", "gg::k_like", nlik)
    line = {"metric": "trial-likelihoods/s, likelihood-only pass (BASELINE config 3)", "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": Wm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS["c3"][3], "trials": n, "nchain": nchain,
                       "l2": "warm: back-to-back launches over the same trials (the sweep point's input is 10 B per trial)",
                       "timing": "CUDA events around K back-to-back launches of the likelihood kernel", "commit": git_head()},
            "gpu_launches": K, "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak > 0 else None,
                         "traffic": traffic, "traffic_source": traffic_source, "kernel": "gg::k_like", "flop_per_trial_lik": F_TRIAL[2],
                         "trial_lik_per_launch": nlik, "launch_ms": ms, "kernel_share_of_step": 1.0,
                         "peak_source": "DFMA microbenchmark on this GPU (ggdmc_b200_measure_fp64_tflops), burst"},
            "cpu_baseline": None, "e2e": e2e}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
