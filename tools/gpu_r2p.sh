#!/bin/bash
# round 2, pass P (2 GPUs): tensor parallelism after the GEMM tail work -- TP parity check, the 2-GPU test, bench tp2 (+ dp line)
N=${N:-2}
mkdir -p gpurun_out
S=gpurun_out/r2p_n${N}_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2p_n${N}_$name.out 2> gpurun_out/r2p_n${N}_$name.err
  echo "$name exit=$?" >> $S; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run tp_check 300 $TR --master-port 29541 tools/tp_check.py
[ "$N" = "2" ] && run tp_test 400 python -m pytest -q -p no:cacheprovider -x tests/test_tp_gpu.py -m gpu
run bench_tp 400 $TR --master-port 29542 bench.py --gpus $N --steps 64 --warmup 4
run bench_ref 300 $TR --master-port 29545 bench.py --impl reference --gpus $N --steps 8 --warmup 1
cat $S
for f in tp_check tp_test bench_tp bench_ref; do
  [ -f gpurun_out/r2p_n${N}_$f.out ] || continue
  echo "== $f"; grep "^{" gpurun_out/r2p_n${N}_$f.out | head -2 | cut -c1-2600; tail -n 3 gpurun_out/r2p_n${N}_$f.out | grep -v "^{" | cut -c1-300; grep -v "^\[W\|^W1\|frame #\|OMP_NUM\|\*\*\*\*" gpurun_out/r2p_n${N}_$f.err | tail -n 4 | cut -c1-300
done
exit 0
