"""BASELINE config 3: likelihood-only throughput sweep, 2-accumulator LBA, 1e3 .. 1e7 trials x 15 chains, one GPU.
Prints trial-likelihoods/s of the likelihood kernel alone (CUDA events around back-to-back launches)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ggdmc_b200 import _lib as B, engine as E, synth, workloads as W
from ggdmc_b200.model import Trials

ct, node_1, p_vector, prior = W.sweep_model()
rng = np.random.default_rng(20260103)
base = synth.simulate_subject(ct, node_1, p_vector, 100_000, rng)
nchain = 15
theta = p_vector * (1.0 + 0.05 * rng.uniform(-1, 1, size=(1, nchain, 5)))
for n in (1_000, 10_000, 100_000, 1_000_000, 10_000_000):
    tr = Trials(base.rt[:: 100_000 // n][:n].copy(), base.cell[:: 100_000 // n][:n].copy()) if n <= 100_000 else \
        Trials(np.tile(base.rt, n // 100_000), np.tile(base.cell, n // 100_000))
    order = np.argsort(tr.cell, kind="stable")
    tr = Trials(tr.rt[order], tr.cell[order])
    ll = E.sumloglike(ct, [tr], theta[0][None])[0]
    lp = E.sumlogprior(prior, theta[0])
    tun = E.Tuning(nmc=2, nchain=nchain, thin=1 << 30, nparameter=5, seeds=[1])
    eng = E.Engine(ct, [tr], prior, None, tun, None, [E.PopState(theta, lp[None], ll[None])])
    ms, nlik = eng.time_likelihood(20)
    print(f"{n:>10d} trials x {nchain} chains: {ms * 1e3:9.1f} us per launch, {nlik / (ms * 1e-3):.3e} trial-likelihoods/s "
          f"({513 * nlik / (ms * 1e-3) / 1e12:.2f} TFLOP/s algorithmic)")
    eng.close()
