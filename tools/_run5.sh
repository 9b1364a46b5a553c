mkdir -p gpurun_out; rm -f gpurun_out/timeline2.txt
for hot in 0 40; do
  for fl in 0 64 16; do
    python tools/timeline.py --m 32 --n 28672 --k 4096 --flags $fl --hot $hot --brief >> gpurun_out/timeline2.txt 2>&1
  done
  python tools/timeline.py --m 32 --n 6144 --k 4096 --flags 0 --hot $hot --brief >> gpurun_out/timeline2.txt 2>&1
  python tools/timeline.py --m 32 --n 6144 --k 4096 --flags 64 --hot $hot --brief >> gpurun_out/timeline2.txt 2>&1
done
cat gpurun_out/timeline2.txt
