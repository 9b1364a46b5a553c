#!/bin/bash
# Swap each tools/variants/lib_<name>.so in for the product library (on the GPU box's scratch copy) and time the
# W4A16 GEMM at the Llama-3-8B decode shapes. usage: run_variants.sh name1 name2 ...  (TESTS=1 also runs the W4A16 parity tests)
mkdir -p gpurun_out
cp lite_llama_b200/libb200lite.so /tmp/lib_product.so
for v in "$@"; do
  cp tools/variants/lib_$v.so lite_llama_b200/libb200lite.so
  echo "== $v" | tee -a gpurun_out/variants.txt
  if [ -n "$TESTS" ]; then
    timeout -s KILL 300 python -m pytest -q -p no:cacheprovider -x --timeout 60 tests/test_kernels_gpu.py -m gpu -k "w4a16 or bs1" 2>&1 | tail -2 | tee -a gpurun_out/variants.txt
  fi
  timeout -s KILL 200 python tools/qgemm_probe.py --shapes qkv,o,gate_up,down --flags packed 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  m=%d %-8s %6.2f us  %5.0f GB/s' % (d['m'], d['shape'], d['packed'], d['packed_gbs']))" | tee -a gpurun_out/variants.txt
  if [ -n "$M1" ]; then
  timeout -s KILL 200 python tools/qgemm_probe.py --m 1 --shapes o,gate_up --flags packed 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  m=%d %-8s %6.2f us  %5.0f GB/s' % (d['m'], d['shape'], d['packed'], d['packed_gbs']))" | tee -a gpurun_out/variants.txt
  fi
done
cp /tmp/lib_product.so lite_llama_b200/libb200lite.so
