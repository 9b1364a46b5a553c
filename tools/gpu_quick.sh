#!/bin/bash
# short, tightly bounded check of the quantised GEMMs: parity tests, then a short bench
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 150 python -m pytest -q -p no:cacheprovider --timeout 40 --timeout-method thread -x tests/test_kernels_gpu.py -k "w4a16 or w8a16 or fp8" > gpurun_out/pytest_gemm.log 2>&1
rc=$?; echo "gemm exit=$rc" >> gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gemm.log >> gpurun_out/summary.txt
if [ $rc -eq 0 ]; then
  timeout 60 python tools/timeline.py > gpurun_out/timeline_w4a16.txt 2>&1
  timeout 200 python bench.py --steps 32 --warmup 4 --cpu-layers 1 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench exit=$?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
