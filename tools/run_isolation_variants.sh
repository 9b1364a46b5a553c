#!/bin/bash
# which build of the quantised GEMM passes the back-to-back isolation test (tests/test_kernels_gpu.py)?
cp lite_llama_b200/libb200lite.so /tmp/lib_product.so
T='python -m pytest -q -p no:cacheprovider --timeout 300 tests/test_kernels_gpu.py -m gpu -k back_to_back'
export B2L_TEST_REPS=${B2L_TEST_REPS:-100}
for v in "$@"; do
  echo "== $v"
  case $v in
    product) cp /tmp/lib_product.so lite_llama_b200/libb200lite.so; timeout -s KILL 400 $T 2>&1 | tail -n 4;;
    nostreamk) cp /tmp/lib_product.so lite_llama_b200/libb200lite.so; B2L_STREAMK=0 timeout -s KILL 400 $T 2>&1 | tail -n 4;;
    nopdl) cp /tmp/lib_product.so lite_llama_b200/libb200lite.so; B2L_PDL=0 timeout -s KILL 400 $T 2>&1 | tail -n 4;;
    *) cp tools/variants/lib_$v.so lite_llama_b200/libb200lite.so; timeout -s KILL 400 $T 2>&1 | tail -n 4;;
  esac
done
cp /tmp/lib_product.so lite_llama_b200/libb200lite.so
