#!/bin/bash
# timeline experiments of the quantised GEMM (debug flags: 1 no epilogue wait, 2 no tcgen05.st, 4 no tcgen05.ld, 8 no xsum)
mkdir -p gpurun_out
for f in 0 1 2 4 8 15; do
  timeout 120 python tools/timeline.py --flags $f > gpurun_out/tl_f$f.txt 2>&1
done
grep -H "total cycles" gpurun_out/tl_f*.txt
