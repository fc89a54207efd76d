#!/usr/bin/env bash
# visit r02x: packed fp32 arithmetic in every epilogue mode of the pair GEMM + the training-leg fusions: the whole -m gpu suite,
# same-visit A/B of forward and training step against the library of the previous commit (ab/libtim_b200_old.so), launch list of a training step
set -u
OUT=gpurun_out
TAG=${1:-r02x}
mkdir -p $OUT
rm -f $OUT/grad_parity.json $OUT/forward_parity.json
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rf > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "passed|failed" $OUT/pytest_gpu_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_gpu_$TAG.log | cut -c1-300 | head -n 30
cp tim_b200/libtim_b200.so /tmp/libtim_new.so
for P in new old new old; do
  if [ $P = old ]; then cp ab/libtim_b200_old.so tim_b200/libtim_b200.so; else cp /tmp/libtim_new.so tim_b200/libtim_b200.so; fi
  timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 > $OUT/bench_${P}_$TAG.json 2> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_${P}_$TAG.json"))
r = d["roofline"]
print("lib=$P ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 4), "path", round(r["path_frac"], 4),
      {k: round(v["ms_per_step"], 3) for k, v in r["by_gemm_kind"].items()}, {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["max_rel_l2_vs_oracle"])
PY
  timeout 300 python bench.py --train-only --steps 8 > $OUT/bench_train_${P}_$TAG.json 2>> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_train_${P}_$TAG.json"))
t = d.get("train", d)
print("train lib=$P ms/step", round(t["ms_per_step"], 2), {k: round(v, 2) for k, v in t["breakdown_ms"].items()}, {k: round(v, 2) for k, v in t["class_ms_per_step"].items()}, "path_frac", round(t["path_frac"], 4))
PY
done
cp /tmp/libtim_new.so tim_b200/libtim_b200.so
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name-base demangled -s 2600 -c 320 --csv \
    --log-file $OUT/launches_train_$TAG.csv python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_list_train_$TAG.log 2>&1
echo "ncu train list exit $?"; wc -l $OUT/launches_train_$TAG.csv
timeout 300 python tools/sustained_gemm.py --seconds 3 --out $OUT/sustained_gemm_$TAG.json > $OUT/sustained_gemm_$TAG.log 2>&1
echo "sustained gemm exit $?"; cat $OUT/sustained_gemm_$TAG.log | cut -c1-260
