#!/bin/bash
# round 2, pass F (1 GPU): fusions (rope+kv, combine fold, swiglu epilogue), GEMM early issue + deferred fix-up, one-tile FA fix
mkdir -p gpurun_out
S=gpurun_out/r2f_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2f_$name.out 2> gpurun_out/r2f_$name.err
  echo "$name exit=$?" >> $S; }
run fa_tests 60 python -m pytest -q -p no:cacheprovider -x --timeout 30 tests/test_kernels_gpu.py -m gpu -k "flash_attention"
run new_tests 120 python -m pytest -q -p no:cacheprovider -x --timeout 40 tests/test_kernels_gpu.py tests/test_host_gpu.py -m gpu -k "swiglu or rope or flash_decoding or w4a16 or moe or prefill_then_decode or greedy or programmatic"
run tests 200 python -m pytest -q -p no:cacheprovider --timeout 40 tests -m gpu --deselect tests/test_stress_gpu.py
run stress 400 python -m pytest -q -p no:cacheprovider -x tests/test_stress_gpu.py -m gpu
run bench 200 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
run bench_bs1 200 python bench.py --batch 1 --steps 64 --warmup 4 --no-cpu-baseline --no-breakdown
run bench_nofuse 200 env B2L_FUSE_SWIGLU=0 B2L_FD_FOLD=0 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --no-breakdown
cat $S
for f in fa_tests new_tests tests stress; do echo "== $f"; tail -12 gpurun_out/r2f_$f.out; done
python - <<'PY'
import json
for f in ["bench","bench_bs1","bench_nofuse"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2f_{f}.out") if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], "prefill", round(d["prefill"]["ms"],1), round(d["prefill"]["tflops"]))
        for k in d.get("kernels",[]): print("   ", k["kernel"], round(k["us_per_launch"],1), round(k["frac_hbm"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
