#!/bin/bash
# round 2, pass G (1 GPU): A/B of the swiglu epilogue, MoE with two CTAs per SM, launch list + ncu captures
mkdir -p gpurun_out
S=gpurun_out/r2g_summary.txt; rm -f $S
run() { local name=$1 t=$2; shift 2
  timeout -s KILL $t "$@" > gpurun_out/r2g_$name.out 2> gpurun_out/r2g_$name.err
  echo "$name exit=$?" >> $S; }
run quick_tests 150 python -m pytest -q -p no:cacheprovider -x --timeout 40 tests/test_kernels_gpu.py tests/test_host_gpu.py -m gpu -k "swiglu or moe or w4a16 or prefill_then_decode or flash_decoding or bs1 or llama3"
run bench 200 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
run bench_noswiglu 200 env B2L_FUSE_SWIGLU=0 python bench.py --steps 64 --warmup 4 --no-cpu-baseline
run bench_moe 300 python bench.py --model qwen3-30b-a3b --quant fp8 --steps 32 --warmup 4 --no-cpu-baseline --watchdog-s 280
run triton_moe 200 python tools/bench_triton_ref.py --only fused_moe --out gpurun_out/r2g_triton_moe
NCU="ncu --set full --clock-control none --import-source on -f"
run launches 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-breakdown
run ncu_gateup 200 $NCU -k regex:gemm_quant_kernel -s 2 -c 1 -o gpurun_out/r2g_w4a16_gateup python tools/prof_kernels.py --op w4a16 --m 32 --n 28672 --k 4096
run ncu_o 200 $NCU -k regex:gemm_quant_kernel -s 2 -c 1 -o gpurun_out/r2g_w4a16_o python tools/prof_kernels.py --op w4a16 --m 32 --n 4096 --k 4096
run ncu_fd256 200 $NCU -k regex:flash_decode_split -s 2 -c 1 -o gpurun_out/r2g_flash_decoding_ctx256 python tools/prof_kernels.py --op decode --ctx 256
run ncu_fd2048 200 $NCU -k regex:flash_decode_split -s 2 -c 1 -o gpurun_out/r2g_flash_decoding_ctx2048 python tools/prof_kernels.py --op decode --ctx 2048
run ncu_fa2048 200 $NCU -k regex:fa_tc_kernel -s 2 -c 1 -o gpurun_out/r2g_fa_tc_s2048 python tools/prof_kernels.py --op prefill --m 4 --ctx 2048
run ncu_fa128 200 $NCU -k regex:fa_tc_one_tile -s 2 -c 1 -o gpurun_out/r2g_fa_tc_s128 python tools/prof_kernels.py --op prefill --m 32 --ctx 128
run ncu_moe 200 $NCU -k regex:moe_gemm_kernel -s 4 -c 2 -o gpurun_out/r2g_moe_gemm_t32 python tools/prof_kernels.py --op moe --m 32
run ncu_w8a8 200 $NCU -k regex:gemm_swapab_kernel -s 2 -c 1 -o gpurun_out/r2g_w8a8_gateup_m256 python tools/prof_kernels.py --op w8a8 --m 256 --n 37888 --k 3584
cat $S
tail -8 gpurun_out/r2g_quick_tests.out
python - <<'PY'
import json
for f in ["bench","bench_noswiglu","bench_moe"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2g_{f}.out") if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], "prefill", round(d["prefill"]["ms"],1), round(d["prefill"]["tflops"]))
        for k in d.get("kernels",[]): print("   ", k["kernel"], round(k["us_per_launch"],1), round(k["frac_hbm"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/r2g_triton_moe.md
ls -la gpurun_out/*.ncu-rep
