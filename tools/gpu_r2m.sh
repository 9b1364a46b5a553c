#!/bin/bash
# round 2, pass M (1 GPU): after the tail/set-up fixes of the quantised GEMM -- parity tests, probe, per-CTA stamps, bench
mkdir -p gpurun_out
S=gpurun_out/r2m_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2m_$name.out 2> gpurun_out/r2m_$name.err
  echo "$name exit=$?" >> $S; }
run tests 400 python -m pytest -q -p no:cacheprovider -x --timeout 60 tests/test_kernels_gpu.py tests/test_host_gpu.py -m gpu -k "w4a16 or w8a16 or bs1 or quant or prefill_then_decode or greedy or layer"
run probe 200 python tools/qgemm_probe.py --shapes qkv,o,gate_up,down --flags packed
run probe_m1 200 python tools/qgemm_probe.py --m 1 --shapes qkv,o,gate_up,down --flags packed
run bench 200 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
run smoke 200 python __graft_entry__.py smoke
run stress 500 python -m pytest -q -p no:cacheprovider -x tests/test_stress_gpu.py -m gpu
cp lite_llama_b200/libb200lite.so /tmp/lib_product.so
cp tools/variants/lib_tlite.so lite_llama_b200/libb200lite.so
run tlite 200 python tools/tlite.py
cp /tmp/lib_product.so lite_llama_b200/libb200lite.so
cat $S
tail -3 gpurun_out/r2m_tests.out; tail -3 gpurun_out/r2m_stress.out
cat gpurun_out/r2m_probe.out gpurun_out/r2m_probe_m1.out | cut -c1-200
tail -2 gpurun_out/r2m_smoke.out gpurun_out/r2m_smoke.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2m_bench.out") if l.startswith("{")][-1])
print("bench", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]))
for k in d.get("kernels",[]): print("   ", k["kernel"], round(k["us_per_launch"],1), round(k["frac_hbm"],3))
PY
grep -A22 "^== o \|^== gate_up" gpurun_out/r2m_tlite.out
