#!/bin/bash
# round 2, pass B (2 GPUs): tensor parallelism -- op checks, TP=2 vs TP=1 logits, graph loop, latency; bench at tp2
mkdir -p gpurun_out
S=gpurun_out/r2b_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2b_$name.out 2> gpurun_out/r2b_$name.err
  echo "$name exit=$?" >> $S; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2b_topo.txt 2>&1
run tp_check 400 $TR --master-port 29541 tools/tp_check.py
run bench_tp2 400 $TR --master-port 29542 bench.py --gpus 2 --steps 64 --warmup 4
run bench_tp2_nccl 400 env B2L_TP_NVLINK=0 $TR --master-port 29543 bench.py --gpus 2 --steps 64 --warmup 4 --no-breakdown
run bench_dp2 400 $TR --master-port 29544 bench.py --gpus 2 --steps 64 --warmup 4 --parallel dp --no-breakdown
run bench_ref2 300 $TR --master-port 29545 bench.py --impl reference --gpus 2 --steps 8 --warmup 1
cat $S
for f in tp_check bench_tp2 bench_tp2_nccl bench_dp2 bench_ref2; do echo "== $f"; grep "^{" gpurun_out/r2b_$f.out | cut -c1-2500; tail -c 1500 gpurun_out/r2b_$f.err; done
