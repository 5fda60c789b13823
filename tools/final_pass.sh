#!/bin/sh
# The round's closing measurements on one B200 (gpurun -- sh tools/final_pass.sh): the GPU test suite, smoke(), one bench
# line per BASELINE config, the sustained C4 line, the persistent-kernel opt-in and the reference arm.  Everything goes
# to gpurun_out/final/; what is judged is copied into profiles/ by hand.
O=gpurun_out/final
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
for wl in c4 c1 c2 c3 c5; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 > $O/bench_line_$wl.json 2> $O/bench_line_$wl.err
done
timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --sustain-s 5 --no-cpu-baseline > $O/bench_line_c4_sustained.json 2> $O/bench_line_c4_sustained.err
GGDMC_B200_PERSIST=1 timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_line_c4_persistent.json 2> $O/bench_line_c4_persistent.err
GGDMC_B200_NO_PERSIST=1 timeout 600 python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_line_c1_launches.json 2> $O/bench_line_c1_launches.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_line.json 2> $O/bench_reference_line.err
tail -3 $O/pytest_gpu.log; cat $O/smoke.log | tail -2
for f in $O/bench_line_*.json; do echo $f; python tools/benchline.py < $f; done
tail -c 600 $O/bench_reference_line.json
