#!/usr/bin/env bash
# visit r02w: training-leg fusions (linear1 + GELU with two outputs = GEMM mode 8, linear2-dgrad + GELU' = mode 9): the training tests,
# same-visit A/B of the training step (TIM_B200_TRAIN_FUSE=1 / 0), launch list of a fused training step
set -u
OUT=gpurun_out
TAG=${1:-r02w}
mkdir -p $OUT
rm -f $OUT/grad_parity.json
timeout 1200 python -m pytest tests/test_gpu_train.py -q -rf --no-header -p no:cacheprovider > $OUT/pytest_train_$TAG.log 2>&1
echo "train tests exit $?"; grep -E "passed|failed" $OUT/pytest_train_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_train_$TAG.log | cut -c1-300 | head -n 20
for P in 1 0 1 0; do
  TIM_B200_TRAIN_FUSE=$P timeout 300 python bench.py --train-only --steps 8 > $OUT/bench_train_fuse${P}_$TAG.json 2>> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_train_fuse${P}_$TAG.json"))
t = d.get("train", d)
print("train FUSE=$P ms/step", round(t["ms_per_step"], 2), {k: round(v, 2) for k, v in t["breakdown_ms"].items()}, {k: round(v, 2) for k, v in t["class_ms_per_step"].items()}, "path_frac", round(t["path_frac"], 4))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name-base demangled -s 2600 -c 320 --csv \
    --log-file $OUT/launches_train_$TAG.csv python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_list_train_$TAG.log 2>&1
echo "ncu train list exit $?"; wc -l $OUT/launches_train_$TAG.csv
