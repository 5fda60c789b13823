#!/bin/sh
# One point of the strong-scaling table and the multi-GPU parity log on an N-GPU box: gpurun --gpus N -- sh tools/scaling_pass.sh N [parity]
N=$1
O=gpurun_out/scale
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n$N.json 2> $O/bench_n$N.err
grep "^{" $O/bench_n$N.json | python tools/benchline.py
if [ "${2:-}" = parity ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tests/multi_gpu_worker.py > $O/multi_gpu_parity_${N}ranks.log 2>&1
  echo "parity exit $?" >> $O/multi_gpu_parity_${N}ranks.log
  tail -8 $O/multi_gpu_parity_${N}ranks.log
fi
