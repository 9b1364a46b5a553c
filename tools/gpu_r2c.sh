#!/bin/bash
# round 2, pass C (1 GPU): the reference Triton kernels and the reference engine on the same B200
mkdir -p gpurun_out
S=gpurun_out/r2c_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2c_$name.out 2> gpurun_out/r2c_$name.err
  echo "$name exit=$?" >> $S; }
run ab_tiny 300 python tools/ref_engine_ab.py --model tiny-llama --layers 2 --quant int4 --batch 4,1 --prompt-len 16 --gen 24 --continuous 6 --out gpurun_out/r2c_ab_tiny.json
run ab_tiny_fp16 300 python tools/ref_engine_ab.py --model tiny-llama --layers 2 --quant fp16 --batch 4 --prompt-len 16 --gen 24 --out gpurun_out/r2c_ab_tiny_fp16.json
run triton_ops 700 python tools/bench_triton_ref.py --out gpurun_out/r2c_triton_vs_b200 --budget-s 500
run ab_llama 900 python tools/ref_engine_ab.py --model llama-3-8b --layers 32 --quant int4 --batch 32,1 --prompt-len 128 --gen 64 --out gpurun_out/r2c_ab_llama3_8b_w4a16.json
cat $S
for f in ab_tiny ab_tiny_fp16 ab_llama; do echo "== $f"; grep "^{" gpurun_out/r2c_$f.out | cut -c1-3500; tail -c 1200 gpurun_out/r2c_$f.err; done
echo "== triton_ops"; cat gpurun_out/r2c_triton_vs_b200.md; tail -c 800 gpurun_out/r2c_triton_ops.err
