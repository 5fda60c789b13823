"""ggdmc_b200: B200-native (CUDA, sm_100a) engine for ggdmc's DE-MCMC / LBA hot path.

The product is `libggdmc_b200.so` (C ABI in include/ggdmc_b200.h).  This package is its host-side mirror of the
reference's interface over ctypes:

    ggdmc_b200.api      run_subject / run_hyper / run on objects with the reference's S4 slot names
    ggdmc_b200.init     initialise_theta / initialise_phi (candidates scored in bulk on the GPU)
    ggdmc_b200.engine   the C ABI one to one (arrays in, arrays out) and the resident Engine
    ggdmc_b200.model    flattening of model / data / prior objects;  ggdmc_b200.rda  reader for .rda fixtures

Nothing here computes a density or a proposal; without a CUDA device every compute call raises (no CPU fallback).
"""
__version__ = "0.1.0"
__all__ = ["api", "engine", "init", "model", "rda", "synth", "workloads"]
