#!/usr/bin/env bash
# visit r03g: bias column sums of the backward on a side stream (under the weight-gradient GEMMs); RNG-state parity of the drop-in:
# the whole -m gpu suite, same-visit A/B of the training step TIM_B200_TRAIN_SIDE=1 / 0
set -u
OUT=gpurun_out
TAG=${1:-r03g}
mkdir -p $OUT
rm -f $OUT/grad_parity.json $OUT/forward_parity.json
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rf > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "passed|failed" $OUT/pytest_gpu_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_gpu_$TAG.log | cut -c1-400 | head -n 30
for P in 1 0 1 0; do
  TIM_B200_TRAIN_SIDE=$P timeout 300 python bench.py --train-only --steps 8 > $OUT/bench_train_side${P}_$TAG.json 2>> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_train_side${P}_$TAG.json"))
t = d.get("train", d)
print("train SIDE=$P ms/step", round(t["ms_per_step"], 2), {k: round(v, 2) for k, v in t["breakdown_ms"].items()}, {k: round(v, 2) for k, v in t["class_ms_per_step"].items()}, "path_frac", round(t["path_frac"], 4))
PY
done
TIM_B200_TRAIN_SIDE=1 timeout 300 python bench.py --workload cfg4 --train-only --steps 6 > $OUT/bench_train_cfg4_side1_$TAG.json 2>> $OUT/bench_$TAG.err
TIM_B200_TRAIN_SIDE=0 timeout 300 python bench.py --workload cfg4 --train-only --steps 6 > $OUT/bench_train_cfg4_side0_$TAG.json 2>> $OUT/bench_$TAG.err
python - <<PY
import json
for p in (1, 0):
    d = json.load(open(f"$OUT/bench_train_cfg4_side{p}_$TAG.json")); t = d.get("train", d)
    print("cfg4 train SIDE=", p, round(t["ms_per_step"], 2), {k: round(v, 2) for k, v in t["breakdown_ms"].items()})
PY
