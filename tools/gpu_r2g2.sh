#!/bin/bash
# round 2, pass G2 (1 GPU): attention v2 (S in registers, O in TMEM), MoE block rule, launch list of a decode step
mkdir -p gpurun_out
S=gpurun_out/r2g2_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2g2_$name.out 2> gpurun_out/r2g2_$name.err
  echo "$name exit=$?" >> $S; }
run fa_tests 60 python -m pytest -q -p no:cacheprovider -x --timeout 30 tests/test_kernels_gpu.py -m gpu -k "flash_attention"
run tests 240 python -m pytest -q -p no:cacheprovider --timeout 60 tests -m gpu --deselect tests/test_stress_gpu.py
run triton_fa 200 python tools/bench_triton_ref.py --only flash_attention --out gpurun_out/r2g2_triton_fa
run triton_moe 200 python tools/bench_triton_ref.py --only fused_moe --out gpurun_out/r2g2_triton_moe
run bench_moe 300 python bench.py --model qwen3-30b-a3b --quant fp8 --steps 32 --warmup 4 --no-cpu-baseline --watchdog-s 280
run launches 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --layers 8 --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown
run ncu_fa2048 200 ncu --set full --clock-control none --import-source on -f -k regex:fa_tc_kernel -s 2 -c 1 -o gpurun_out/r2g2_fa_tc_v2_s2048 python tools/prof_kernels.py --op prefill --m 4 --ctx 2048
cat $S
echo "== fa_tests"; tail -12 gpurun_out/r2g2_fa_tests.out
echo "== tests"; tail -12 gpurun_out/r2g2_tests.out
cat gpurun_out/r2g2_triton_fa.md gpurun_out/r2g2_triton_moe.md
python - <<'PY'
import json
for f in ["bench_moe"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2g2_{f}.out") if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), "prefill", round(d["prefill"]["ms"],1), round(d["prefill"]["tflops"]))
        for k in d.get("kernels",[]): print("   ", k["kernel"], round(k["us_per_launch"],1), round(k["frac_hbm"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
wc -l gpurun_out/launches.csv
