#!/bin/bash
# quick A/B pass: parity tests of the quantised GEMMs, probe, per-CTA stamps
mkdir -p gpurun_out
S=gpurun_out/r2n_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2n_$name.out 2> gpurun_out/r2n_$name.err
  echo "$name exit=$?" >> $S; }
run tests 400 python -m pytest -q -p no:cacheprovider -x --timeout 60 tests/test_kernels_gpu.py -m gpu -k "w4a16 or w8a16 or bs1 or quant"
run probe 200 python tools/qgemm_probe.py --shapes qkv,o,gate_up,down --flags packed
cp lite_llama_b200/libb200lite.so /tmp/lib_product.so
cp tools/variants/lib_tlite.so lite_llama_b200/libb200lite.so
run tlite 200 python tools/tlite.py --shapes o:4096x4096,down:4096x14336,gate_up:28672x4096
cp /tmp/lib_product.so lite_llama_b200/libb200lite.so
[ -n "$BENCH" ] && run bench 200 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
[ -n "$STRESS" ] && run stress 500 python -m pytest -q -p no:cacheprovider -x tests/test_stress_gpu.py -m gpu
[ -n "$STRESS" ] && run host 400 python -m pytest -q -p no:cacheprovider -x --timeout 120 tests/test_host_gpu.py -m gpu
cat $S
tail -n 3 gpurun_out/r2n_tests.out
cat gpurun_out/r2n_probe.out | cut -c1-200
[ -n "$STRESS" ] && tail -n 3 gpurun_out/r2n_stress.out gpurun_out/r2n_host.out
[ -n "$BENCH" ] && grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r2n_bench.out | head -4
cat gpurun_out/r2n_tlite.out
