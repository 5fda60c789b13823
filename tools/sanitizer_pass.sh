#!/bin/sh
# compute-sanitizer over the sampler's trajectory tests (gpurun -- sh tools/sanitizer_pass.sh): memcheck on every kernel path,
# racecheck (shared-memory hazards) on the hierarchical trajectories -> gpurun_out/sanitizer.txt
O=gpurun_out/sanitizer.txt
: > $O
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sampler.py -m gpu -q -x -k "trajectory or execution_paths" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|at 0x|by thread" | head -40 >> $O
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_sampler.py -m gpu -q -x -k "hierarchical_trajectory" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|hazard|at 0x" | head -40 >> $O
cat $O
