#!/usr/bin/env bash
# training-leg visit: gradient parity tests, the training-step bench record of cfg2 (and cfg4 with "cfg4"), optional ncu of the attention backward
set -u
OUT=gpurun_out
TAG=${1:-train}
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_real_reference_gpu.py -q -rf --no-header -p no:cacheprovider -x > $OUT/pytest_train_$TAG.log 2>&1
echo "training tests exit $?"; grep -E "passed|failed" $OUT/pytest_train_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_train_$TAG.log | cut -c1-260 | head -n 20
timeout 600 python bench.py --train-only --steps 8 --warmup 3 > $OUT/bench_train_$TAG.json 2> $OUT/bench_train_$TAG.err
python - <<PY
import json
try:
    t = json.load(open("$OUT/bench_train_$TAG.json"))["train"]
    print("train ms/step", round(t["ms_per_step"], 2), {k: round(v, 2) for k, v in t["breakdown_ms"].items()}, "path_frac", round(t["path_frac"], 4))
    print({k: round(v, 2) for k, v in t["class_ms_per_step"].items() if v > 0})
except Exception as e:
    print("parse failed", e); print(open("$OUT/bench_train_$TAG.err").read()[-2000:])
PY
if [ "${2:-}" = "cfg4" ] || [ "${3:-}" = "cfg4" ]; then
timeout 600 python bench.py --train-only --workload cfg4 --clips 48 --steps 5 --warmup 3 > $OUT/bench_train_cfg4_$TAG.json 2>> $OUT/bench_train_$TAG.err
python - <<PY
import json
try:
    t = json.load(open("$OUT/bench_train_cfg4_$TAG.json"))["train"]
    print("cfg4 train ms/step", round(t["ms_per_step"], 2), {k: round(v, 2) for k, v in t["class_ms_per_step"].items() if v > 0})
except Exception as e:
    print("parse failed", e)
PY
fi
if [ "${2:-}" = "ncu" ] || [ "${3:-}" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attention_bwd_umma" -s 3 -c 1 \
    -o $OUT/prof_attnbwd_$TAG -f python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_attnbwd_$TAG.log 2>&1
echo "ncu attention bwd exit $?"; ls -la $OUT/prof_attnbwd_$TAG.ncu-rep
fi
