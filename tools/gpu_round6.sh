#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  timeout $t python -m pytest -q -p no:cacheprovider --timeout 60 --timeout-method thread "$@" > gpurun_out/pytest_$name.log 2>&1
  echo "$name exit=$?" >> gpurun_out/summary.txt; tail -4 gpurun_out/pytest_$name.log >> gpurun_out/summary.txt; }
run gemm 300 tests/test_kernels_gpu.py -k "w4a16 or w8a16 or fp8"
timeout 120 python tools/timeline.py > gpurun_out/timeline_w4a16.txt 2>&1
timeout 120 python tools/timeline.py --op w8a16 > gpurun_out/timeline_w8a16.txt 2>&1
timeout 600 python tools/microbench.py --which gemm > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench exit=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 32 --warmup 4 --cpu-layers 1 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench exit=$?" >> gpurun_out/summary.txt
tail -c 600 gpurun_out/bench_short.err >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
