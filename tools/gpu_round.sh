#!/usr/bin/env bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of the same command and one full capture of
# the dominant kernel. Everything lands in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/smi_$TAG.txt 2>&1

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
  echo "pytest exit $?" | tee -a $OUT/pytest_gpu_$TAG.log
  tail -n 5 $OUT/pytest_gpu_$TAG.log
fi

timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; cat $OUT/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
cat $OUT/bench_ref_$TAG.json

if [ "${SKIP_NCU:-0}" != "1" ]; then
  # launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv \
      --log-file $OUT/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_$TAG.log 2>&1
  echo "ncu list exit $?"
  # one full capture of the dominant kernel (3 launches from the middle of the run)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_umma -s 40 -c 3 \
      -o $OUT/prof_gemm_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
  echo "ncu full exit $?"
fi

# detection post-processing kernels (soft-NMS of an evaluation-sized input against the reference's compiled extension)
if [ "${SKIP_NMS:-0}" != "1" ]; then
  timeout 200 python tools/nms_bench.py --out $OUT/nms_bench_$TAG.json > $OUT/nms_bench_$TAG.log 2>&1
  echo "nms bench exit $?"; tail -n 1 $OUT/nms_bench_$TAG.log | cut -c1-400
fi
# per-launch summaries of the full captures (what profiles/*_ncu_full_*.csv hold)
for rep in $OUT/prof_gemm_$TAG.ncu-rep; do
  [ -f "$rep" ] && python tools/ncu_summary.py "$rep" > "${rep%.ncu-rep}.csv" 2>/dev/null
done
# secondary baseline: the reference's own PyTorch calls, eager, on the same GPU
if [ "${SKIP_EAGER:-0}" != "1" ]; then
  timeout 300 python tools/eager_baseline.py --out $OUT/eager_baseline_$TAG.json > $OUT/eager_baseline_$TAG.log 2>&1
  echo "eager baseline exit $?"; tail -n 1 $OUT/eager_baseline_$TAG.log | cut -c1-600
fi
