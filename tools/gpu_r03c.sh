#!/usr/bin/env bash
# visit r03c: the other named configurations at HEAD on one GPU: cfg3 (Perception, E = 1536, head_dim 192), cfg4 as the headline workload, cfg2 with bf16 operands
set -u
OUT=gpurun_out
TAG=${1:-r03c}
mkdir -p $OUT
for W in cfg3 cfg4; do
  timeout 400 python bench.py --workload $W --no-extras --no-cpu-baseline --steps 20 > $OUT/bench_${W}_$TAG.json 2> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_${W}_$TAG.json"))
r = d["roofline"]
print("$W ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 4), "path", round(r["path_frac"], 4),
      {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["max_rel_l2_vs_oracle"], d["config"]["clips_per_gpu_per_step"])
PY
done
timeout 400 python bench.py --dtype bf16 --no-extras --no-cpu-baseline --steps 20 > $OUT/bench_cfg2_bf16_$TAG.json 2>> $OUT/bench_$TAG.err
python - <<PY
import json
d = json.load(open("$OUT/bench_cfg2_bf16_$TAG.json"))
r = d["roofline"]
print("cfg2 bf16 ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 4), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"])
PY
