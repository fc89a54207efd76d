#!/usr/bin/env bash
# visit r03b: HEAD with programmatic dependent launch: the whole -m gpu suite (incl. the CUDA-graph capture test), smoke, the DEFAULT bench line, the reference arm, launch list of a bench step
set -u
OUT=gpurun_out
TAG=${1:-r03b}
mkdir -p $OUT
rm -f $OUT/grad_parity.json $OUT/forward_parity.json
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rf > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "passed|failed" $OUT/pytest_gpu_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_gpu_$TAG.log | cut -c1-300 | head -n 30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 3 $OUT/smoke_$TAG.log
T0=$(date +%s)
timeout 1200 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $? wall $(( $(date +%s) - T0 )) s"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$TAG.json"))
    print("fwd ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "fp32io", d["e2e_fp32_io"]["value"], "fp32 logits", d["e2e_fp32_logits"]["value"], "bank", d["e2e_resident_bank"]["value"])
    print("gemm frac", d["roofline"]["frac"], "path frac", d["roofline"]["path_frac"], d["roofline"]["class_ms_per_step"], "clocks", d["clocks"])
    t = d["train"]; print("train ms/step", t["ms_per_step"], t["breakdown_ms"], "path_frac", t["path_frac"], t["class_ms_per_step"])
    c = d["cfg4"]; print("cfg4 fwd ms", c["ms_per_step"], "value", c["value"], "e2e", c["e2e"]["value"], "train ms", c["train"]["ms_per_step"])
    print("sweep", [(r["S"], r["clips_per_gpu"], round(r["ms_per_step"], 1), round(r["path_frac"], 3)) for r in d["sweep_cfg5"]["rows"]])
    print("eager", {k: (round(v["ms_per_step"], 1), round(v["tim_b200_speedup"], 2)) for k, v in d["gpu_eager_baseline"]["modes"].items()})
    print("gemm_vs_cublas", d.get("gemm_vs_cublas"))
    print("cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-3000:])
PY
T0=$(date +%s)
timeout 900 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
echo "reference arm wall $(( $(date +%s) - T0 )) s"
python -c "
import json; d=json.load(open('$OUT/bench_ref_$TAG.json')); print('reference arm', d['value'], d['steps'], d['warmup'], d['spread'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name-base demangled -s 400 -c 140 --csv \
    --log-file $OUT/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launch list exit $?"; wc -l $OUT/launches_bench_$TAG.csv
