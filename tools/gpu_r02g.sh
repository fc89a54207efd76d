#!/usr/bin/env bash
# visit r02g: deeper-pipeline attention forward (attention_umma3.cu): kernel tests, isolated timing v1 / v2 / v3, same-visit bench A/B, ncu
set -u
OUT=gpurun_out
TAG=${1:-r02g}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -rf --no-header -p no:cacheprovider -k "attention_kernel or golden or named_configs or fuzz or full_size" > $OUT/pytest_attn3_$TAG.log 2>&1
echo "forward tests (attention v3) exit $?"; grep -E "passed|failed" $OUT/pytest_attn3_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_attn3_$TAG.log | cut -c1-260 | head -n 20
timeout 300 python tools/attn_bench.py --shapes cfg2,cfg3,cfg4,hd64 > $OUT/attn_bench_$TAG.txt 2>&1; cat $OUT/attn_bench_$TAG.txt
for S in 2 3 4; do TIM_B200_ATTN_STAGES=$S timeout 120 python tools/attn_bench.py --versions 3 --shapes cfg2 2>&1 | grep "v3" | sed "s/^/stages=$S /"; done | tee -a $OUT/attn_bench_$TAG.txt
for PF in 0 1 2 4; do TIM_B200_ATTN_PF=2,$PF timeout 120 python tools/attn_bench.py --versions 3 --shapes cfg2 2>&1 | grep "v3 B" | sed "s/^/pf=$PF /"; done | tee -a $OUT/attn_bench_$TAG.txt
for A in 3 2 3 2; do
  TIM_B200_ATTN=$A timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 > $OUT/bench_attn${A}_$TAG.json 2> $OUT/bench_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_attn${A}_$TAG.json")); r = d["roofline"]
    print("ATTN=$A ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 4), "path", round(r["path_frac"], 4),
          {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["max_rel_l2_vs_oracle"])
except Exception as e:
    print("parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-2000:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attention_umma3" -s 3 -c 2 \
    -o $OUT/prof_attn3_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_attn3_$TAG.log 2>&1
echo "ncu attn3 exit $?"
[ -f $OUT/prof_attn3_$TAG.ncu-rep ] && python tools/ncu_summary.py $OUT/prof_attn3_$TAG.ncu-rep > $OUT/prof_attn3_$TAG.csv 2>/dev/null && grep -E "time_duration|dram__bytes|tensor_cycles|warps_active|issue_active" $OUT/prof_attn3_$TAG.csv
