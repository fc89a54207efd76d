#!/usr/bin/env bash
# attention forward A/B: kernel tests, isolated timing, in-step bench (TIM_B200_ATTN=4 vs 2), optional ncu of the v4 kernel
set -u
OUT=gpurun_out
TAG=${1:-attn}
mkdir -p $OUT
timeout 300 python tools/attn_bench.py --versions 1,2,4 --shapes cfg2,cfg4,hd64,small > $OUT/attn_bench_$TAG.txt 2>&1; cat $OUT/attn_bench_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -q -rf --no-header -p no:cacheprovider -x -k "attention or golden or named_configs or fuzz or full_size or dropout" > $OUT/pytest_attn4_$TAG.log 2>&1
echo "tests (attention v4 default) exit $?"; grep -E "passed|failed" $OUT/pytest_attn4_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_attn4_$TAG.log | cut -c1-260 | head -n 20
for A in 4 2 4; do
  TIM_B200_ATTN=$A timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 > $OUT/bench_attn${A}_$TAG.json 2> $OUT/bench_$TAG.err
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
if [ "${2:-}" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attention_umma4" -s 2 -c 1 \
    -o $OUT/prof_attn4_$TAG -f python tools/attn_bench.py --versions 4 --shapes cfg2 --iters 3 > $OUT/ncu_attn4_$TAG.log 2>&1
echo "ncu attn4 exit $?"; ls -la $OUT/prof_attn4_$TAG.ncu-rep
fi
