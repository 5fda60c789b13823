#!/bin/sh
# The ncu evidence of a round, in one go on the GPU box (gpurun -- sh tools/ncu_profiles.sh <commit>): writes into gpurun_out/r02_*;
# the summaries that are judged are then copied into profiles/ by hand.  Numbers printed by runs under ncu
# are never bench values.
set -u
C=${1:-unknown}
PART=${2:-all}   # gpurun copies at most 64 MiB back: "lists" = steps 1-2 and the k_like capture, "sampler" = the two sampler captures
O=gpurun_out
mkdir -p $O
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv"
if [ "$PART" != sampler ]; then
# 1. launch list of the default bench command (every kernel with its device time)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_launches.log 2>&1
# 2. DRAM traffic of the dominant kernel of every workload
ncu $M -k regex:k_like -s 20 -c 8 --log-file $O/r02_traffic_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu $M -k regex:k_like -s 6 -c 8 --log-file $O/r02_traffic_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu $M -k regex:k_like -s 20 -c 8 --log-file $O/r02_traffic_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu $M -k regex:k_like -s 4 -c 4 --log-file $O/r02_traffic_c3.csv python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu $M -k regex:k_sampler -s 3 -c 4 --log-file $O/r02_traffic_c1.csv python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
# 3. full captures: the likelihood kernel on C4 (launch path), the persistent sampler kernel on C1 and, opt-in, on C4
ncu --set full --clock-control none --import-source on -k regex:k_like -s 24 -c 1 -o $O/r02_k_like_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
fi
if [ "$PART" != lists ]; then
ncu --set full --clock-control none --import-source on -k regex:k_sampler -s 4 -c 1 -o $O/r02_k_sampler_c1 python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
GGDMC_B200_PERSIST=1 ncu --set full --clock-control none --import-source on -k regex:k_sampler -s 4 -c 1 -o $O/r02_k_sampler_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
fi
ls -la $O | grep r02_
