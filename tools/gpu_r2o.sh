#!/bin/bash
# round 2, pass O (1 GPU): full validation after the GEMM tail work -- whole GPU suite (incl. stress), smoke, bench lines,
# per-CTA stamps, launch list + ncu captures of the two ends of the W4A16 GEMM range
mkdir -p gpurun_out
S=gpurun_out/r2o_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2o_$name.out 2> gpurun_out/r2o_$name.err
  echo "$name exit=$?" >> $S; }
run tests 900 python -m pytest -q -p no:cacheprovider -x --timeout 120 tests -m gpu
run smoke 200 python __graft_entry__.py smoke
run bench 200 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
run bench_bs1 200 python bench.py --batch 1 --steps 64 --warmup 4 --no-cpu-baseline
run bench_fp16 300 python bench.py --quant fp16 --steps 64 --warmup 4 --no-cpu-baseline
run bench_w8a8 300 python bench.py --model qwen2.5-7b --quant smoothquant --batch 256 --steps 32 --warmup 4 --no-cpu-baseline --no-breakdown
run bench_moe 300 python bench.py --model qwen3-30b-a3b --quant fp8 --steps 32 --warmup 4 --no-cpu-baseline --watchdog-s 280
NCU="ncu --set full --clock-control none --import-source on -f"
run launches 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --layers 8 --steps 2 --warmup 1 --no-cpu-baseline --no-breakdown
run ncu_gateup 200 $NCU -k regex:gemm_quant_kernel -s 2 -c 1 -o gpurun_out/r2o_w4a16_gateup python tools/prof_kernels.py --op w4a16 --m 32 --n 28672 --k 4096
run ncu_o 200 $NCU -k regex:gemm_quant_kernel -s 2 -c 1 -o gpurun_out/r2o_w4a16_o python tools/prof_kernels.py --op w4a16 --m 32 --n 4096 --k 4096
cp lite_llama_b200/libb200lite.so /tmp/lib_product.so
cp tools/variants/lib_tlite.so lite_llama_b200/libb200lite.so
run tlite 200 python tools/tlite.py
cp /tmp/lib_product.so lite_llama_b200/libb200lite.so
cat $S
tail -n 4 gpurun_out/r2o_tests.out
tail -n 2 gpurun_out/r2o_smoke.out gpurun_out/r2o_smoke.err
python - <<'PY'
import json
for f in ["bench","bench_bs1","bench_fp16","bench_w8a8","bench_moe"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2o_{f}.out") if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], "prefill", round(d["prefill"]["ms"],1), round(d["prefill"]["tflops"]), "step frac", round(d["step_roofline"]["frac"],3))
        for k in d.get("kernels",[]): print("   ", k["kernel"], round(k["us_per_launch"],1), round(k["frac_hbm"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la gpurun_out/r2o_*.ncu-rep
