#!/bin/bash
# First hardware pass: parity tests in isolated processes (a hang in one group must not hide the rest),
# then per-kernel timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, args...
  local name=$1; shift; local t=$1; shift
  timeout $t python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider --timeout 120 "$@" > gpurun_out/pytest_$name.log 2>&1
  echo "$name exit=$?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/pytest_$name.log >> gpurun_out/summary.txt
}
rm -f gpurun_out/summary.txt
run elementwise 400 -k "rmsnorm or swiglu or rope or kv_ops or update_kv"
run decode 600 -k "flash_decoding"
run prefill 400 -k "flash_attention"
run linear 400 -k "linear"
run w8a8 300 -k "smoothquant"
run w4a16 400 -k "w4a16"
run w8a16 300 -k "w8a16 or fp8"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
