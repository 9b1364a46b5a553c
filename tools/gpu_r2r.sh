#!/bin/bash
# round 2, pass R (1 GPU): fused norm + int8 activation quantisation (SmoothQuant decoder) -- whole suite + W8A8 lines
mkdir -p gpurun_out
S=gpurun_out/r2r_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2r_$name.out 2> gpurun_out/r2r_$name.err
  echo "$name exit=$?" >> $S; }
run tests 600 python -m pytest -q -p no:cacheprovider -x --timeout 120 tests -m gpu --deselect tests/test_stress_gpu.py
run bench_w8a8 200 python bench.py --model qwen2.5-7b --quant smoothquant --batch 256 --steps 32 --warmup 4 --no-cpu-baseline --no-breakdown
run bench_w8a8_unfused 200 env B2L_FUSE_NORM_QUANT=0 python bench.py --model qwen2.5-7b --quant smoothquant --batch 256 --steps 32 --warmup 4 --no-cpu-baseline --no-breakdown
cat $S
tail -n 6 gpurun_out/r2r_tests.out
python - <<'PY'
import json
for f in ["bench_w8a8","bench_w8a8_unfused"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2r_{f}.out") if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"])
    except Exception as e:
        print(f, "ERR", e)
PY
