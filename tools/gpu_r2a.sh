#!/bin/bash
# round 2, pass A: hang root cause + stress + tests + bench at batch 1 and 32 (every step bounded)
mkdir -p gpurun_out
S=gpurun_out/r2a_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2a_$name.out 2> gpurun_out/r2a_$name.err
  echo "$name exit=$?" >> $S; }
B2L_HANG_MS=1500 run hang_legacy 120 python tools/hang_repro.py
run hang_fixed 120 python tools/hang_repro.py --fixed
run gemm_m1 60 python tools/qgemm_probe.py --m 1 --flags base,packed
run stress 900 python -m pytest -q -p no:cacheprovider -x tests/test_stress_gpu.py -m gpu
run tests 600 python -m pytest -q -p no:cacheprovider -x tests -m gpu --deselect tests/test_stress_gpu.py
run bench_bs1 300 python bench.py --batch 1 --steps 64 --warmup 4 --no-cpu-baseline
run bench 400 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
cat $S
for f in hang_legacy hang_fixed gemm_m1 bench_bs1 bench; do echo "== $f"; tail -c 3000 gpurun_out/r2a_$f.out; tail -c 600 gpurun_out/r2a_$f.err; done
echo "== stress"; tail -15 gpurun_out/r2a_stress.out
echo "== tests"; tail -8 gpurun_out/r2a_tests.out
