for PF in 2,1 2,2 2,3 1,1 0,0; do
  TIM_B200_ATTN_PF=$PF timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_pf.json")); r = d["roofline"]
print("PF=$PF ms/step", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in r["class_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"])
PY
done
