#!/bin/bash
# bounded stress of the quantised GEMM under back-to-back launches (tight timeouts: a hang costs seconds, not minutes)
mkdir -p gpurun_out
S=gpurun_out/stress.txt; rm -f $S
timeout 40 python tools/timeline.py --m 32 --n 28672 --k 4096 --hot 40 --packed --brief > gpurun_out/st_tl_gu.txt 2>&1; echo "tl gate_up packed exit=$?" >> $S
timeout 40 python tools/timeline.py --m 32 --n 28672 --k 4096 --hot 40 --brief > gpurun_out/st_tl_gu_ref.txt 2>&1; echo "tl gate_up ref exit=$?" >> $S
timeout 40 python tools/timeline.py --m 32 --n 4096 --k 14336 --hot 40 --packed --brief > gpurun_out/st_tl_dn.txt 2>&1; echo "tl down packed exit=$?" >> $S
timeout 60 python tools/qgemm_probe.py --shapes down,gate_up --flags base,packed > gpurun_out/st_probe.jsonl 2>&1; echo "probe down,gate_up exit=$?" >> $S
timeout 60 python tools/qgemm_probe.py --shapes qkv,o --flags base,packed >> gpurun_out/st_probe.jsonl 2>&1; echo "probe qkv,o exit=$?" >> $S
cat $S gpurun_out/st_probe.jsonl
