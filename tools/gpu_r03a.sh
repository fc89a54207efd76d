#!/usr/bin/env bash
# visit r03a (PDL levels 2 / 1 / 0): programmatic dependent launch (PDL) of the pair GEMM and the tcgen05 attention forward: forward + training tests, A/B TIM_B200_PDL=1 / 0
set -u
OUT=gpurun_out
TAG=${1:-r03a}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --no-header -p no:cacheprovider -rf > $OUT/pytest_fwd_$TAG.log 2>&1
echo "forward tests exit $?"; grep -E "passed|failed" $OUT/pytest_fwd_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_fwd_$TAG.log | cut -c1-300 | head -n 10
timeout 600 python -m pytest tests/test_gpu_train.py -q -x --no-header -p no:cacheprovider -rf  > $OUT/pytest_train_$TAG.log 2>&1
echo "train tests exit $?"; grep -E "passed|failed" $OUT/pytest_train_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_train_$TAG.log | cut -c1-300 | head -n 10
for P in 2 1 0 2 1 0; do
  TIM_B200_PDL=$P timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 > $OUT/bench_pdl${P}_$TAG.json 2> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_pdl${P}_$TAG.json"))
r = d["roofline"]
print("PDL=$P ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "fp32io", round(d["e2e_fp32_io"]["value"]), "gemm frac", round(r["frac"], 4),
      {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["max_rel_l2_vs_oracle"])
PY
done
for P in 2 0; do
  TIM_B200_PDL=$P timeout 300 python bench.py --train-only --steps 8 > $OUT/bench_train_pdl${P}_$TAG.json 2>> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_train_pdl${P}_$TAG.json"))
t = d.get("train", d)
print("train PDL=$P ms/step", round(t["ms_per_step"], 2), {k: round(v, 2) for k, v in t["breakdown_ms"].items()})
PY
done
