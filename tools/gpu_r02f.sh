#!/usr/bin/env bash
# visit r02f: the whole -m gpu suite at HEAD, the full bench line, the reference arm, ncu evidence for the remaining HBM-bound kernels
set -u
OUT=gpurun_out
TAG=${1:-r02f}
mkdir -p $OUT
rm -f $OUT/grad_parity.json $OUT/forward_parity.json
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest -m gpu exit $?"; tail -n 4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 3 $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$TAG.json"))
    print("fwd ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "fp32io", d["e2e_fp32_io"]["value"], "fp16 logits", d["e2e_fp32_logits"]["value"], "bank", d["e2e_resident_bank"]["value"])
    print("gemm frac", d["roofline"]["frac"], "path frac", d["roofline"]["path_frac"], d["roofline"]["class_ms_per_step"], "clocks", d["clocks"])
    t = d["train"]; print("train ms/step", t["ms_per_step"], t["breakdown_ms"], "path_frac", t["path_frac"], t["class_ms_per_step"])
    c = d["cfg4"]; print("cfg4 fwd ms", c["ms_per_step"], "value", c["value"], "e2e", c["e2e"]["value"], "train ms", c["train"]["ms_per_step"])
    print("sweep", [(r["S"], r["clips_per_gpu"], round(r["ms_per_step"], 1), round(r["path_frac"], 3)) for r in d["sweep_cfg5"]["rows"]])
    print("eager", {k: (round(v["ms_per_step"], 1), round(v["tim_b200_speedup"], 2)) for k, v in d["gpu_eager_baseline"]["modes"].items()})
    print("cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-3000:])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
python -c "
import json; d=json.load(open('$OUT/bench_ref_$TAG.json')); print('reference arm', d['value'], d['steps'], d['warmup'], d['spread'])"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gather_rows_kernel|cast_kernel<__half>" -s 330 -c 6 \
    -o $OUT/prof_castgather_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --clips 512 > $OUT/ncu_castgather_$TAG.log 2>&1
echo "ncu cast / gather exit $?"
[ -f $OUT/prof_castgather_$TAG.ncu-rep ] && python tools/ncu_summary.py $OUT/prof_castgather_$TAG.ncu-rep > $OUT/prof_castgather_$TAG.csv 2>/dev/null && cut -c1-260 $OUT/prof_castgather_$TAG.csv | head -n 12
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attention_bwd_umma" -s 3 -c 1 \
    -o $OUT/prof_attnbwd_umma_$TAG -f python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_attnbwd_umma_$TAG.log 2>&1
[ -f $OUT/prof_attnbwd_umma_$TAG.ncu-rep ] && python tools/ncu_summary.py $OUT/prof_attnbwd_umma_$TAG.ncu-rep > $OUT/prof_attnbwd_umma_$TAG.csv 2>/dev/null && grep -E "time_duration|dram__bytes|tensor_cycles" $OUT/prof_attnbwd_umma_$TAG.csv
