#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/timeline.py --m 32 --n 28672 --k 4096 --hot 40 --packed > gpurun_out/tl_gateup.txt 2>&1
timeout 100 python tools/timeline.py --m 32 --n 28672 --k 4096 --hot 40 --packed --flags 16 > gpurun_out/tl_gateup_raw.txt 2>&1
tail -50 gpurun_out/tl_gateup.txt
