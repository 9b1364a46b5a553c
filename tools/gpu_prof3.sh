#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:gemm_quant -s 2 -c 1 -f -o gpurun_out/prof_w4a16_v6 python tools/prof_kernels.py --op w4a16 --n 28672 --k 4096 > gpurun_out/prof_w4a16.log 2>&1; echo "w4a16 $?"
