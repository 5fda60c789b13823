mkdir -p gpurun_out
for ns in 0 1 2 3 4 6; do
  echo "== c2 nsplit $ns"; if [ $ns = 0 ]; then unset GGDMC_B200_NSPLIT; else export GGDMC_B200_NSPLIT=$ns; fi
  timeout 200 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python tools/benchline.py
done
for ns in 0 1 2 3 4; do
  echo "== c4-128 nsplit $ns"; if [ $ns = 0 ]; then unset GGDMC_B200_NSPLIT; else export GGDMC_B200_NSPLIT=$ns; fi
  timeout 200 python bench.py --workload c4 --subjects 128 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python tools/benchline.py
done
for ns in 1 2; do
  echo "== c4 nsplit $ns"; export GGDMC_B200_NSPLIT=$ns
  timeout 200 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python tools/benchline.py
done
for ns in 1 2 3; do
  echo "== c4-256 nsplit $ns"; export GGDMC_B200_NSPLIT=$ns
  timeout 200 python bench.py --workload c4 --subjects 256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python tools/benchline.py
done
