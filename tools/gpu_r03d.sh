#!/usr/bin/env bash
# visit r03d: packed fp32 arithmetic in the attention forward (exp2 / row-sum pass, epilogue): kernel and golden tests, isolated timing, same-visit A/B
set -u
OUT=gpurun_out
TAG=${1:-r03d}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q --no-header -p no:cacheprovider -rf -k "attention or golden or named_configs or graph" > $OUT/pytest_fwd_$TAG.log 2>&1
echo "forward tests exit $?"; grep -E "passed|failed" $OUT/pytest_fwd_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_fwd_$TAG.log | cut -c1-300 | head -n 10
timeout 600 python -m pytest tests/test_gpu_train.py -q --no-header -p no:cacheprovider -rf -k "golden or dropout" > $OUT/pytest_train_$TAG.log 2>&1
echo "train tests (dropout variants of the forward kernel) exit $?"; grep -E "passed|failed" $OUT/pytest_train_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_train_$TAG.log | cut -c1-300 | head -n 10
cp tim_b200/libtim_b200.so /tmp/libtim_new.so
for P in new old new old; do
  if [ $P = old ]; then cp ab/libtim_b200_old.so tim_b200/libtim_b200.so; else cp /tmp/libtim_new.so tim_b200/libtim_b200.so; fi
  timeout 200 python tools/attn_bench.py --versions 4 --shapes cfg2,cfg4,hd64 > $OUT/attn_bench_${P}_$TAG.txt 2>&1; echo "attn_bench $P:"; grep -E "v4|ms" $OUT/attn_bench_${P}_$TAG.txt | cut -c1-200 | head -6
  timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 > $OUT/bench_${P}_$TAG.json 2> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_${P}_$TAG.json"))
r = d["roofline"]
print("lib=$P ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 4),
      {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["max_rel_l2_vs_oracle"])
PY
done
cp /tmp/libtim_new.so tim_b200/libtim_b200.so
