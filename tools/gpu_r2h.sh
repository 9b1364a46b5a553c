#!/bin/bash
# round 2, pass H (N GPUs, default 2): tensor parallelism with the packet exchange kernel; TP vs DP lines; MoE under TP
N=${N:-2}
mkdir -p gpurun_out
S=gpurun_out/r2h_n${N}_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2h_n${N}_$name.out 2> gpurun_out/r2h_n${N}_$name.err
  echo "$name exit=$?" >> $S; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  run quick 120 python -m pytest -q -p no:cacheprovider -x --timeout 40 tests/test_kernels_gpu.py tests/test_host_gpu.py -m gpu -k "flash_attention or rmsnorm or moe or llama3 or bs1"
  run triton_fa 200 python tools/bench_triton_ref.py --only flash_attention --out gpurun_out/r2h_triton_fa
  run tp_check 300 $TR --master-port 29541 tools/tp_check.py
  run tp_test 400 python -m pytest -q -p no:cacheprovider -x tests/test_tp_gpu.py -m gpu
fi
if [ "$N" = "8" ]; then
  run tp_check 300 $TR --master-port 29541 tools/tp_check.py
fi
if [ "$N" = "4" ]; then
  run quick 120 python -m pytest -q -p no:cacheprovider -x --timeout 40 tests/test_kernels_gpu.py -m gpu -k "fused_moe or argmax or rmsnorm"
fi
run bench_tp 400 $TR --master-port 29542 bench.py --gpus $N --steps 64 --warmup 4
run bench_ref 300 $TR --master-port 29545 bench.py --impl reference --gpus $N --steps 8 --warmup 1
if [ "$N" != "8" ]; then
  run bench_moe_tp 500 $TR --master-port 29546 bench.py --gpus $N --model qwen3-30b-a3b --quant fp8 --steps 32 --warmup 4 --no-dp-line --watchdog-s 450
fi
cat $S
for f in quick tp_check tp_test bench_tp bench_ref bench_moe_tp; do
  [ -f gpurun_out/r2h_n${N}_$f.out ] || continue
  echo "== $f"; grep "^{" gpurun_out/r2h_n${N}_$f.out | cut -c1-2200; tail -4 gpurun_out/r2h_n${N}_$f.out | grep -v "^{" | cut -c1-300; grep -v "^\[W\|^W1\|frame #\|OMP_NUM\|\*\*\*\*" gpurun_out/r2h_n${N}_$f.err | tail -6 | cut -c1-300
done

[ -f gpurun_out/r2h_triton_fa.md ] && cat gpurun_out/r2h_triton_fa.md
