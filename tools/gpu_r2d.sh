#!/bin/bash
# round 2, pass D (1 GPU): tcgen05 prefill attention + dense tile heuristic: tests, then the Triton comparison rows
mkdir -p gpurun_out
S=gpurun_out/r2d_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2d_$name.out 2> gpurun_out/r2d_$name.err
  echo "$name exit=$?" >> $S; }
run fa_tests 300 python -m pytest -q -p no:cacheprovider -x tests/test_kernels_gpu.py -m gpu -k "flash_attention"
run tests 600 python -m pytest -q -p no:cacheprovider tests -m gpu --deselect tests/test_stress_gpu.py
run triton_fa 300 python tools/bench_triton_ref.py --only flash_attention --out gpurun_out/r2d_triton_fa
run triton_sq 300 python tools/bench_triton_ref.py --only smoothquant --out gpurun_out/r2d_triton_sq
run bench 300 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
cat $S
echo "== fa_tests"; tail -30 gpurun_out/r2d_fa_tests.out
echo "== tests"; tail -15 gpurun_out/r2d_tests.out
cat gpurun_out/r2d_triton_fa.md gpurun_out/r2d_triton_sq.md
grep "^{" gpurun_out/r2d_bench.out | cut -c1-600; tail -5 gpurun_out/r2d_bench.err
