#!/usr/bin/env bash
# visit r02d: two-plane residual stream (mode 7): forward parity, same-visit A/B against mode 5, ncu of the producer GEMMs
set -u
OUT=gpurun_out
TAG=${1:-r02d}
mkdir -p $OUT
rm -f $OUT/forward_parity.json
timeout 900 python -m pytest tests/test_gpu_parity.py -q -rf --no-header -p no:cacheprovider > $OUT/pytest_fwd_$TAG.log 2>&1
echo "forward tests exit $?"; grep -E "passed|failed" $OUT/pytest_fwd_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_fwd_$TAG.log | cut -c1-300 | head -n 20
timeout 600 python -m pytest tests/test_gpu_train.py -q -rf --no-header -p no:cacheprovider -k "attention_bwd or named_configs or dropin" > $OUT/pytest_train_$TAG.log 2>&1
echo "train tests exit $?"; grep -E "passed|failed" $OUT/pytest_train_$TAG.log | tail -n 2
for P in 1 0 1 0; do
  TIM_B200_PLANES=$P timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 > $OUT/bench_planes${P}_$TAG.json 2> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_planes${P}_$TAG.json"))
r = d["roofline"]
print("PLANES=$P ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e_fp32", round(d["e2e_fp32_io"]["value"]), "gemm frac", round(r["frac"], 4), "path", round(r["path_frac"], 4),
      {k: round(v["ms_per_step"], 3) for k, v in r["by_gemm_kind"].items()}, {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["max_rel_l2_vs_oracle"])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:linear_umma2_kernel" -s 44 -c 9 \
    -o $OUT/prof_gemm_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_gemm_$TAG.log 2>&1
echo "ncu gemm exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --kernel-name-base demangled -s 300 -c 140 --csv --log-file $OUT/launches_fwd_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_list_fwd_$TAG.log 2>&1
echo "ncu fwd list exit $?"
[ -f $OUT/prof_gemm_$TAG.ncu-rep ] && python tools/ncu_summary.py $OUT/prof_gemm_$TAG.ncu-rep > $OUT/prof_gemm_$TAG.csv 2>/dev/null
