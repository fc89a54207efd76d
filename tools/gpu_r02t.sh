#!/usr/bin/env bash
# visit r02t: A/B of the within-tile residual L2 prefetch (TIM_B200_RES_PF=2 vs 0), ncu --set full of the mode-7 producer GEMMs (out_proj, linear2),
# compute-sanitizer at HEAD over the real-width (head_dim 128: tcgen05 attention forward / backward) cfg1 case
set -u
OUT=gpurun_out
TAG=${1:-r02t}
mkdir -p $OUT
for P in 2 0 2 0; do
  TIM_B200_RES_PF=$P timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 > $OUT/bench_respf${P}_$TAG.json 2> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_respf${P}_$TAG.json"))
r = d["roofline"]
print("RES_PF=$P ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 4), "path", round(r["path_frac"], 4),
      {k: round(v["ms_per_step"], 3) for k, v in r["by_gemm_kind"].items()}, {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["max_rel_l2_vs_oracle"])
PY
done
TIM_B200_RES_PF=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:linear_umma2_kernel<__half, 7" -s 24 -c 2 \
    -o $OUT/prof_gemm7_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_gemm7_$TAG.log 2>&1
echo "ncu gemm7 exit $?"; ls -la $OUT/prof_gemm7_$TAG.ncu-rep
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck_smoke_$TAG.log 2>&1
echo "memcheck smoke exit $?"; tail -n 3 $OUT/sanitizer_memcheck_smoke_$TAG.log | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/train_smoke.py recog_cfg1 fp16 > $OUT/sanitizer_memcheck_train_$TAG.log 2>&1
echo "memcheck train (cfg1 widths) exit $?"; tail -n 3 $OUT/sanitizer_memcheck_train_$TAG.log | cut -c1-300
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/train_smoke.py recog_cfg1 fp16 > $OUT/sanitizer_racecheck_train_$TAG.log 2>&1
echo "racecheck train (cfg1 widths) exit $?"; tail -n 3 $OUT/sanitizer_racecheck_train_$TAG.log | cut -c1-300
