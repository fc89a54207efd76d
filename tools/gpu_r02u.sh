#!/usr/bin/env bash
# visit r02u: packed-fp32 (FADD2 / FFMA2) epilogue of the mode-7 producer GEMMs: forward parity, same-visit A/B against the previous library
# (ab/libtim_b200_old.so, built from HEAD's gemm_umma2.cu), ncu --set full of out_proj / linear2, launch list
set -u
OUT=gpurun_out
TAG=${1:-r02u}
mkdir -p $OUT
rm -f $OUT/forward_parity.json
timeout 900 python -m pytest tests/test_gpu_parity.py -q -rf --no-header -p no:cacheprovider > $OUT/pytest_fwd_$TAG.log 2>&1
echo "forward tests exit $?"; grep -E "passed|failed" $OUT/pytest_fwd_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_fwd_$TAG.log | cut -c1-300 | head -n 20
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
done
cp /tmp/libtim_new.so tim_b200/libtim_b200.so
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:linear_umma2_kernel<__half, \(int\)7" -s 24 -c 2 \
    -o $OUT/prof_gemm7_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_gemm7_$TAG.log 2>&1
echo "ncu gemm7 exit $?"; ls -la $OUT/prof_gemm7_$TAG.ncu-rep; tail -n 3 $OUT/ncu_gemm7_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --kernel-name-base demangled -s 300 -c 140 --csv --log-file $OUT/launches_fwd_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_list_fwd_$TAG.log 2>&1
echo "ncu fwd list exit $?"
