#!/usr/bin/env bash
# visit r02y: packed fp32 softmax passes in the tcgen05 attention backward, assembly without the fp32 token copy: the whole -m gpu suite,
# same-visit A/B of the training step against the previous commit's library, the DEFAULT bench line (with the gemm_vs_cublas record)
set -u
OUT=gpurun_out
TAG=${1:-r02y}
mkdir -p $OUT
rm -f $OUT/grad_parity.json $OUT/forward_parity.json
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rf > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "passed|failed" $OUT/pytest_gpu_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_gpu_$TAG.log | cut -c1-300 | head -n 30
cp tim_b200/libtim_b200.so /tmp/libtim_new.so
for P in new old new old; do
  if [ $P = old ]; then cp ab/libtim_b200_old.so tim_b200/libtim_b200.so; else cp /tmp/libtim_new.so tim_b200/libtim_b200.so; fi
  timeout 300 python bench.py --train-only --steps 8 > $OUT/bench_train_${P}_$TAG.json 2>> $OUT/bench_$TAG.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_train_${P}_$TAG.json"))
t = d.get("train", d)
print("train lib=$P ms/step", round(t["ms_per_step"], 2), {k: round(v, 2) for k, v in t["breakdown_ms"].items()}, {k: round(v, 2) for k, v in t["class_ms_per_step"].items()}, "path_frac", round(t["path_frac"], 4))
PY
done
cp /tmp/libtim_new.so tim_b200/libtim_b200.so
T0=$(date +%s)
timeout 1200 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $? wall $(( $(date +%s) - T0 )) s"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$TAG.json"))
    print("fwd ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "fp32io", d["e2e_fp32_io"]["value"], "bank", d["e2e_resident_bank"]["value"])
    print("gemm frac", d["roofline"]["frac"], "path frac", d["roofline"]["path_frac"], d["roofline"]["class_ms_per_step"], "clocks", d["clocks"])
    t = d["train"]; print("train ms/step", t["ms_per_step"], t["breakdown_ms"], "path_frac", t["path_frac"])
    c = d["cfg4"]; print("cfg4 fwd ms", c["ms_per_step"], "value", c["value"], "e2e", c["e2e"]["value"], "train ms", c["train"]["ms_per_step"])
    print("gemm_vs_cublas", d["gemm_vs_cublas"])
    print("cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-3000:])
PY
