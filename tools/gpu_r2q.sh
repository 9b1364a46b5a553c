#!/bin/bash
# round 2, pass Q (1 GPU): after the weight-stage release fix -- whole GPU suite, determinism probe, bench lines
mkdir -p gpurun_out
S=gpurun_out/r2q_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2q_$name.out 2> gpurun_out/r2q_$name.err
  echo "$name exit=$?" >> $S; }
run tests 900 python -m pytest -q -p no:cacheprovider -x --timeout 120 tests -m gpu
run nondet64 100 python tools/nondet_probe.py --m 64 --reps 300
run nondet36 100 python tools/nondet_probe.py --m 36 --reps 300
run nondet256 100 python tools/nondet_probe.py --m 256 --reps 100
run nondet32 100 python tools/nondet_probe.py --m 32 --n 28672 --k 4096 --reps 200
run probe 200 python tools/qgemm_probe.py --shapes qkv,o,gate_up,down --flags packed
run bench 200 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
run bench_bs1 200 python bench.py --batch 1 --steps 64 --warmup 4 --no-cpu-baseline --no-breakdown
run smoke 200 python __graft_entry__.py smoke
cat $S
tail -n 3 gpurun_out/r2q_tests.out
head -n 2 gpurun_out/r2q_nondet*.out | cut -c1-200
cat gpurun_out/r2q_probe.out | cut -c1-110
python - <<'PY'
import json
for f in ["bench","bench_bs1"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2q_{f}.out") if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "prefill", round(d["prefill"]["ms"],1), round(d["prefill"]["tflops"]))
        for k in d.get("kernels",[]): print("   ", k["kernel"], round(k["us_per_launch"],1), round(k["frac_hbm"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 1 gpurun_out/r2q_smoke.out | cut -c1-120
