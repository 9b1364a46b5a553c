#!/bin/bash
# round 2, pass K (1 GPU): steady-state device timelines of the long W4A16 GEMMs + debug-flag A/B on the current build; smoke()
mkdir -p gpurun_out
S=gpurun_out/r2k_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2k_$name.out 2> gpurun_out/r2k_$name.err
  echo "$name exit=$?" >> $S; }
run smoke 200 python __graft_entry__.py smoke
run tl_gateup_cold 100 python tools/timeline.py --m 32 --n 28672 --k 4096 --packed
run tl_gateup_hot 100 python tools/timeline.py --m 32 --n 28672 --k 4096 --hot 40 --packed
run tl_down_hot 100 python tools/timeline.py --m 32 --n 4096 --k 14336 --hot 40 --packed
run probe 300 python tools/qgemm_probe.py --shapes gate_up,down,o --flags packed,packed+dbg_inst,packed+skipall,packed+halfring,packed+noX,packed+few_mma,packed+no_sttm,packed+skip_drain,packed+raw_operand,packed+skipall+noX,packed+skipall+halfring
cat $S
tail -3 gpurun_out/r2k_smoke.out gpurun_out/r2k_smoke.err
cat gpurun_out/r2k_probe.out
