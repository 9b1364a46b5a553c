#!/bin/bash
mkdir -p gpurun_out
for r in 0 1; do
  B2L_DQ_ROLLED=$r timeout 100 python tools/qgemm_probe.py --flags packed > gpurun_out/probe_rolled$r.jsonl 2> gpurun_out/probe_rolled$r.err
  B2L_DQ_ROLLED=$r timeout 60 python tools/timeline.py --m 32 --n 4096 --k 4096 --hot 40 --packed > gpurun_out/tl_o_rolled$r.txt 2>&1
  B2L_DQ_ROLLED=$r timeout 60 python tools/timeline.py --m 32 --n 28672 --k 4096 --hot 40 --packed > gpurun_out/tl_gu_rolled$r.txt 2>&1
done
timeout 60 python -m pytest -q -p no:cacheprovider --timeout 60 -x tests/test_kernels_gpu.py -m gpu -k "packed" 2>&1 | tail -2
B2L_DQ_ROLLED=1 timeout 60 python -m pytest -q -p no:cacheprovider --timeout 60 -x tests/test_kernels_gpu.py -m gpu -k "packed" 2>&1 | tail -2
cat gpurun_out/probe_rolled0.jsonl gpurun_out/probe_rolled1.jsonl
