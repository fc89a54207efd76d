#!/usr/bin/env bash
# visit r02s: index-check test, forward parity at HEAD, same-visit A/B of the residual L2 prefetch in the pair GEMM (TIM_B200_RES_PF), launch list
set -u
OUT=gpurun_out
TAG=${1:-r02s}
mkdir -p $OUT
rm -f $OUT/forward_parity.json
timeout 900 python -m pytest tests/test_gpu_parity.py -q -rf --no-header -p no:cacheprovider > $OUT/pytest_fwd_$TAG.log 2>&1
echo "forward tests exit $?"; grep -E "passed|failed" $OUT/pytest_fwd_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_fwd_$TAG.log | cut -c1-300 | head -n 20
timeout 600 python -m pytest tests/test_gpu_train.py -q -rf --no-header -p no:cacheprovider -k "golden or named_configs or dropin" > $OUT/pytest_train_$TAG.log 2>&1
echo "train tests exit $?"; grep -E "passed|failed" $OUT/pytest_train_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_train_$TAG.log | cut -c1-300 | head -n 20
for P in 1 0 1 0; do
  TIM_B200_RES_PF=$P timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 > $OUT/bench_respf${P}_$TAG.json 2> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_respf${P}_$TAG.json"))
r = d["roofline"]
print("RES_PF=$P ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 4), "path", round(r["path_frac"], 4),
      {k: round(v["ms_per_step"], 3) for k, v in r["by_gemm_kind"].items()}, {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["max_rel_l2_vs_oracle"])
PY
done
for P in 1 0; do
  TIM_B200_RES_PF=$P timeout 300 python bench.py --train-only --steps 6 > $OUT/bench_train_respf${P}_$TAG.json 2>> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_train_respf${P}_$TAG.json"))
t = d.get("train", d)
print("train RES_PF=$P ms/step", t["ms_per_step"], t["breakdown_ms"], {k: round(v, 2) for k, v in t["class_ms_per_step"].items()})
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --kernel-name-base demangled -s 300 -c 140 --csv --log-file $OUT/launches_fwd_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_list_fwd_$TAG.log 2>&1
echo "ncu fwd list exit $?"
