#!/bin/bash
# round 2, pass E (1 GPU): one-tile attention, rope+kv fusion, MoE models; tight time-outs first, then measurements
mkdir -p gpurun_out
S=gpurun_out/r2e_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2e_$name.out 2> gpurun_out/r2e_$name.err
  echo "$name exit=$?" >> $S; }
run new_tests 90 python -m pytest -q -p no:cacheprovider -x --timeout 60 tests/test_kernels_gpu.py tests/test_host_gpu.py -m gpu -k "flash_attention or rope or moe or prefill_then_decode or greedy or programmatic"
run tests 300 python -m pytest -q -p no:cacheprovider --timeout 60 tests -m gpu --deselect tests/test_stress_gpu.py
run triton_fa 200 python tools/bench_triton_ref.py --only flash_attention --out gpurun_out/r2e_triton_fa
run bench 200 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
run bench_moe 300 python bench.py --model qwen3-30b-a3b --quant fp8 --steps 32 --warmup 4 --no-cpu-baseline --watchdog-s 280
run bench_w8a8 300 python bench.py --model qwen2.5-7b --quant smoothquant --batch 256 --steps 32 --warmup 4 --no-cpu-baseline --watchdog-s 280
run ab_cont 200 python tools/ref_engine_ab.py --model tiny-llama --layers 2 --quant int4 --batch 4 --prompt-len 16 --gen 24 --continuous 6 --arms C --out gpurun_out/r2e_ab_tiny_continuous.json
cat $S
echo "== new_tests"; tail -25 gpurun_out/r2e_new_tests.out
echo "== tests"; tail -12 gpurun_out/r2e_tests.out
cat gpurun_out/r2e_triton_fa.md
for f in bench bench_moe bench_w8a8 ab_cont; do echo "== $f"; grep "^{" gpurun_out/r2e_$f.out | cut -c1-1800; tail -c 800 gpurun_out/r2e_$f.err; done
