#!/usr/bin/env bash
# visit r02e: tcgen05 attention backward: kernel tests, whole training leg, A/B against the warp-MMA kernels, ncu
set -u
OUT=gpurun_out
TAG=${1:-r02e}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_train.py -q -rf --no-header -p no:cacheprovider -k "attention_bwd" > $OUT/pytest_attnbwd_$TAG.log 2>&1
echo "attention bwd kernel tests exit $?"; grep -E "passed|failed" $OUT/pytest_attnbwd_$TAG.log | tail -n 2; grep -E "^FAILED|rel-L2|Error" $OUT/pytest_attnbwd_$TAG.log | cut -c1-260 | head -n 24
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_real_reference_gpu.py -q -rf --no-header -p no:cacheprovider -k "not attention_bwd and not wgrad" > $OUT/pytest_train_$TAG.log 2>&1
echo "train tests exit $?"; grep -E "passed|failed" $OUT/pytest_train_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_train_$TAG.log | cut -c1-260 | head -n 12
timeout 600 python -m pytest tests/test_gpu_parity.py -q --no-header -p no:cacheprovider -k "golden" > $OUT/pytest_fwd_$TAG.log 2>&1
echo "forward golden tests exit $?"; tail -n 2 $OUT/pytest_fwd_$TAG.log | cut -c1-200
for A in 0 1 0; do
  TIM_B200_ATTN_BWD=$A timeout 400 python bench.py --train-only --steps 6 > $OUT/bench_train_bwd${A}_$TAG.json 2> $OUT/bench_train_$TAG.err
  python - <<PY
import json
try:
    t = json.load(open("$OUT/bench_train_bwd${A}_$TAG.json"))["train"]
    print("ATTN_BWD=$A train ms", round(t["ms_per_step"], 2), {k: round(v, 2) for k, v in t["breakdown_ms"].items()}, {k: round(v, 2) for k, v in t["class_ms_per_step"].items() if v > 0}, "path_frac", round(t["path_frac"], 4))
except Exception as e:
    print("parse failed", e); print(open("$OUT/bench_train_$TAG.err").read()[-2000:])
PY
done
timeout 600 python bench.py --train-only --workload cfg4 --clips 48 --steps 4 > $OUT/bench_train_cfg4_$TAG.json 2>> $OUT/bench_train_$TAG.err
python -c "
import json
t=json.load(open('$OUT/bench_train_cfg4_$TAG.json'))['train']; print('cfg4 train ms', t['ms_per_step'], {k: round(v,2) for k,v in t['class_ms_per_step'].items() if v>0})"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attention_bwd_umma" -s 3 -c 2 \
    -o $OUT/prof_attnbwd_umma_$TAG -f python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_attnbwd_umma_$TAG.log 2>&1
echo "ncu attn bwd umma exit $?"
[ -f $OUT/prof_attnbwd_umma_$TAG.ncu-rep ] && python tools/ncu_summary.py $OUT/prof_attnbwd_umma_$TAG.ncu-rep > $OUT/prof_attnbwd_umma_$TAG.csv 2>/dev/null && cut -c1-200 $OUT/prof_attnbwd_umma_$TAG.csv
