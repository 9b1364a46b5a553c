#!/bin/bash
# quick probe: parity of the quantised GEMMs, device timelines, A/B timing at the Llama-3-8B shapes
mkdir -p gpurun_out
S=gpurun_out/summary.txt; rm -f $S
FLAGS=${FLAGS:-base,packed}
timeout 300 python -m pytest -q -p no:cacheprovider --timeout 60 -x tests/test_kernels_gpu.py -m gpu -k "w4a16 or w8a16 or fp8 or quant" > gpurun_out/pytest_gemm.log 2>&1
echo "gemm tests exit=$?" >> $S; tail -3 gpurun_out/pytest_gemm.log >> $S
timeout 100 python tools/timeline.py --m 32 --n 28672 --k 4096 --hot 40 --packed > gpurun_out/tl_gateup.txt 2>&1
timeout 100 python tools/timeline.py --m 32 --n 4096 --k 4096 --hot 40 --packed > gpurun_out/tl_o.txt 2>&1
timeout 200 python tools/qgemm_probe.py --flags $FLAGS > gpurun_out/probe1.jsonl 2> gpurun_out/probe1.err
echo "probe exit=$?" >> $S
cat gpurun_out/probe1.jsonl >> $S
cat $S
