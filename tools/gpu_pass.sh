#!/bin/bash
# One consolidated hardware pass (GPU minutes are scarce): parity suite, smoke, headline bench, the reference
# arm, the ncu launch list of the bench command and one `--set full` capture of the two dominant kernels.
# Everything lands in gpurun_out/; summaries worth keeping are copied to profiles/ afterwards (tools/keep_profiles.py).
mkdir -p gpurun_out
S=gpurun_out/summary.txt
rm -f $S gpurun_out/timeline.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
STEPS=${STEPS:-"tests smoke bench ref launches ncu_gemm ncu_decode micro"}
for step in $STEPS; do
  t0=$(date +%s)
  case $step in
    tests)
      timeout 420 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
      echo "tests exit=$?" >> $S; tail -4 gpurun_out/pytest_gpu.log >> $S ;;
    smoke)
      timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
      echo "smoke exit=$?" >> $S; tail -2 gpurun_out/smoke.log >> $S ;;
    bench)
      timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "bench exit=$?" >> $S; tail -3 gpurun_out/bench.err >> $S ;;
    bench1)
      timeout 300 python bench.py --batch 1 --no-cpu-baseline > gpurun_out/bench_bs1.json 2> gpurun_out/bench_bs1.err
      echo "bench1 exit=$?" >> $S; tail -3 gpurun_out/bench_bs1.err >> $S ;;
    ref)
      timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
      echo "ref exit=$?" >> $S; tail -3 gpurun_out/bench_ref.err >> $S ;;
    launches)
      timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-breakdown > gpurun_out/launches_bench.log 2>&1
      echo "launches exit=$?" >> $S ;;
    ncu_gemm)
      timeout 240 ncu --set full --clock-control none --import-source on -k regex:gemm_quant_kernel -s 2 -c 1 -f -o gpurun_out/w4a16_gateup \
        python tools/prof_kernels.py --op w4a16 --m 32 --n 28672 --k 4096 > gpurun_out/ncu_gemm.log 2>&1
      echo "ncu_gemm exit=$?" >> $S ;;
    ncu_decode)
      timeout 240 ncu --set full --clock-control none --import-source on -k regex:flash_decode_split -s 2 -c 1 -f -o gpurun_out/flash_decoding_ctx8192 \
        python tools/prof_kernels.py --op decode --ctx 8192 > gpurun_out/ncu_decode.log 2>&1
      echo "ncu_decode exit=$?" >> $S ;;
    timeline)
      for shape in "6144 4096" "4096 4096" "28672 4096" "4096 14336"; do
        set -- $shape
        timeout 120 python tools/timeline.py --m 32 --n $1 --k $2 >> gpurun_out/timeline.txt 2>&1
      done
      echo "timeline exit=$?" >> $S ;;
    bench_nopdl)
      B2L_PDL=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err
      echo "bench_nopdl exit=$?" >> $S; tail -3 gpurun_out/bench_nopdl.err >> $S ;;
    bench_nosk)
      B2L_STREAMK=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_nosk.json 2> gpurun_out/bench_nosk.err
      echo "bench_nosk exit=$?" >> $S; tail -3 gpurun_out/bench_nosk.err >> $S ;;
    probe)
      timeout 300 python tools/qgemm_probe.py > gpurun_out/qgemm_probe.jsonl 2> gpurun_out/qgemm_probe.err
      echo "probe exit=$?" >> $S; tail -3 gpurun_out/qgemm_probe.err >> $S ;;
    micro)
      timeout 300 python tools/microbench.py > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err
      echo "micro exit=$?" >> $S ;;
  esac
  echo "  [$step took $(( $(date +%s) - t0 )) s]" >> $S
done
cat $S
