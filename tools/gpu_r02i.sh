#!/usr/bin/env bash
# visit r02i: the three tests that failed in r02h, attention forward after the DROP template split (isolated + in-step), one
# ncu --set full capture of attention_umma_kernel with source-level sampling (where do the cycles of a tile go)
set -u
OUT=gpurun_out
TAG=${1:-r02i}
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rf -k "patch_model_dropin or gradients_with_dropout or attention_kernel or dropout" > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" $OUT/pytest_gpu_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_gpu_$TAG.log | cut -c1-300 | head -n 30
timeout 300 python tools/attn_bench.py --versions 2 --shapes cfg2,cfg3,cfg4,hd64 > $OUT/attn_bench_$TAG.txt 2>&1; cat $OUT/attn_bench_$TAG.txt
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$TAG.json")); r = d["roofline"]
    print("ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 4), "path", round(r["path_frac"], 4),
          {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-2000:])
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attention_umma_kernel" -s 2 -c 1 \
    -o $OUT/prof_attn2_$TAG -f python tools/attn_bench.py --versions 2 --shapes cfg2 --iters 3 > $OUT/ncu_attn2_$TAG.log 2>&1
echo "ncu attn2 exit $?"; ls -la $OUT/prof_attn2_$TAG.ncu-rep
