"""DDM ("fastdm") likelihood throughput on one GPU: k_like_ddm alone (CUDA events around back-to-back launches) for the
three variability regimes, the algorithmic work per trial counted by the oracle (series evaluations and terms), and the
reference's own object code (likelihood_class::ddm_likelihood of src/de.o, -O0) and the -O2 oracle timed on one host core
over a bounded sample of the same inputs."""
import os, sys, time
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ggdmc_b200 import engine as E, workloads as W
from ggdmc_b200.model import Trials
from oracle import binding as ob

# nominal costs as in SURVEY.md 8(d): add/mul/cmp 1, fma 2, div 8, sqrt 8, exp 24, log 28, sin 28
F_EVAL_NOVAR, F_EVAL_VAR = 205.0, 231.0   # one series evaluation without its terms: t/a^2, factor, eps, get_N, norm, scaling
F_SMALL, F_LARGE = 37.0, 60.0             # one small-time term (exp + div), one large-time term (exp + sin)
F_TRIAL = 30.0                            # rt - t_offset, log (or its running-product twin), sum

only = sys.argv[1:] or None  # regime names to run (default: all three)
rng = np.random.default_rng(20260105)
peak = E.measure_fp64_tflops()
print(f"FP64 FMA peak measured on this GPU: {peak:.2f} TFLOP/s")
for name, zero, S, reps in (("no variability", ("st0", "sv", "sz"), 256, 10), ("sv", ("st0", "sz"), 256, 10), ("sv+sz+st0", (), 32, 3)):
    if only and name not in only:
        continue
    ct, truth, prior = W.ddm_model(fixed=zero)  # a variability that is off is a constant 0, not a free parameter
    om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar, type=ob.MODEL_DDM)
    nchain = 3 * ct.npar
    pool = W.ddm_simulate(truth, 6000, rng, pnames=ct.pnames)
    subjects = []
    for s_ in range(S):
        idx = np.sort(rng.choice(len(pool.rt), 768, replace=False))
        subjects.append(Trials(pool.rt[idx].copy(), pool.cell[idx].copy()))
    theta = truth[None, None, :] * (1.0 + 0.03 * rng.uniform(-1, 1, size=(S, nchain, ct.npar)))
    ll = E.sumloglike(ct, subjects, theta)
    lp = np.stack([E.sumlogprior(prior, theta[s_]) for s_ in range(S)])
    tun = E.Tuning(nmc=2, nchain=nchain, thin=1 << 30, nparameter=ct.npar, seeds=[1])
    eng = E.Engine(ct, subjects, prior, None, tun, None, [E.PopState(theta[s_][None], lp[s_][None], ll[s_][None]) for s_ in range(S)])
    ms, nlik = eng.time_likelihood(reps)
    eng.close()
    # algorithmic work per trial-likelihood, counted by the oracle on a sample of (subject, chain) pairs
    cnt = (C.c_longlong * 3)()
    ob.lib().orc_ddm_counters(cnt, 1)
    pairs = [(int(rng.integers(S)), int(rng.integers(nchain))) for _ in range(12 if zero else 2)]
    od = {s_: ob.OData(subjects[s_].rt, subjects[s_].cell) for s_, _ in pairs}
    t0 = time.perf_counter()
    for s_, c in pairs:
        ob.sumloglike(om, od[s_], theta[s_, c])
    t_port = time.perf_counter() - t0
    parity = max(abs(ll[s_, c] - ob.sumloglike(om, od[s_], theta[s_, c])) / abs(ll[s_, c]) for s_, c in pairs)
    ob.lib().orc_ddm_counters(cnt, 1)
    n_s = 768 * len(pairs)
    evals, small, large = cnt[0] / n_s, cnt[1] / n_s, cnt[2] / n_s
    flop = F_TRIAL + evals * (F_EVAL_NOVAR if "sv" in zero else F_EVAL_VAR) + small * F_SMALL + large * F_LARGE
    rate = nlik / (ms * 1e-3)
    line = (f"{name:>14s}: {S} subjects x 768 trials x {nchain} chains: {ms:9.3f} ms per launch, {rate:.3e} trial-likelihoods/s; "
            f"per trial {evals:.1f} series evaluations, {small:.1f} small-time + {large:.1f} large-time terms = {flop:.0f} flop "
            f"-> {flop * rate / 1e12:.2f} TFLOP/s algorithmic = {flop * rate / 1e12 / peak:.3f} of peak; "
            f"oracle port (-O2, 1 core): {n_s / t_port:.3e}/s; sums vs oracle: max rel diff {parity:.1e}")
    if ob.ref_lib() is not None:
        ob.ref2_ddm_density(om, od[pairs[0][0]], theta[pairs[0]])  # warm-up: the static ddm_obj is built on the first call
        t0 = time.perf_counter()
        for s_, c in pairs[:max(1, len(pairs) // 3)]:
            ob.ref2_ddm_density(om, od[s_], theta[s_, c])
        t_ref = time.perf_counter() - t0
        line += f"; reference object code (de.o, -O0, 1 core): {768 * max(1, len(pairs) // 3) / t_ref:.3e}/s"
    print(line, flush=True)
