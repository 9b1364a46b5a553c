#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:gemm_quant -s 2 -c 1 -f -o gpurun_out/prof_w4a16_gateup python tools/prof_kernels.py --op w4a16 --n 28672 --k 4096 > gpurun_out/prof_w4a16.log 2>&1; echo "w4a16 $?"
timeout 600 $NCU -k regex:flash_decode_split -s 2 -c 1 -f -o gpurun_out/prof_decode_ctx8192 python tools/prof_kernels.py --op decode --ctx 8192 > gpurun_out/prof_decode.log 2>&1; echo "decode $?"
timeout 600 $NCU -k regex:gemm_swapab -s 2 -c 1 -f -o gpurun_out/prof_dense_lmhead python tools/prof_kernels.py --op dense --n 128256 --k 4096 > gpurun_out/prof_dense.log 2>&1; echo "dense $?"
ls -la gpurun_out/*.ncu-rep
