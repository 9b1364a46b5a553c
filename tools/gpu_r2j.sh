#!/bin/bash
# round 2, pass J (1 GPU): GEMM device timelines, prefill-size GEMM microbench, fp16 decode lines (configs[1]), headline
mkdir -p gpurun_out
S=gpurun_out/r2j_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2j_$name.out 2> gpurun_out/r2j_$name.err
  echo "$name exit=$?" >> $S; }
run quick 150 python -m pytest -q -p no:cacheprovider -x --timeout 40 tests/test_kernels_gpu.py tests/test_host_gpu.py -m gpu -k "rmsnorm or prefill_then_decode or argmax or greedy"
run tl_o 100 python tools/timeline.py --m 32 --n 4096 --k 4096 --hot 40 --packed
run tl_qkv 100 python tools/timeline.py --m 32 --n 6144 --k 4096 --hot 40 --packed
run tl_o_m1 100 python tools/timeline.py --m 1 --n 4096 --k 4096 --hot 40 --packed
run micro_prefill 200 python tools/microbench.py --which prefill_gemm
run bench 200 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
run bench_fp16 300 python bench.py --quant fp16 --steps 64 --warmup 4 --no-cpu-baseline
run bench_fp16_bs1 300 python bench.py --quant fp16 --batch 1 --steps 64 --warmup 4 --no-cpu-baseline --no-breakdown
run bench_bs1 200 python bench.py --batch 1 --steps 64 --warmup 4 --no-cpu-baseline
cat $S
tail -5 gpurun_out/r2j_quick.out
for f in tl_o tl_qkv tl_o_m1; do echo "== $f"; head -40 gpurun_out/r2j_$f.out; done
echo "== micro"; cat gpurun_out/r2j_micro_prefill.out
python - <<'PY'
import json
for f in ["bench","bench_fp16","bench_fp16_bs1","bench_bs1"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2j_{f}.out") if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], "prefill", round(d["prefill"]["ms"],1), round(d["prefill"]["tflops"]), "step frac", round(d["step_roofline"]["frac"],3))
        for k in d.get("kernels",[]): print("   ", k["kernel"], round(k["us_per_launch"],1), round(k["frac_hbm"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
