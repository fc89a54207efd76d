#!/usr/bin/env bash
# visit r02h: the whole -m gpu suite at HEAD (first GPU run of the dropout kernels), smoke, a short bench line
set -u
OUT=gpurun_out
TAG=${1:-r02h}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rf > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "passed|failed" $OUT/pytest_gpu_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_gpu_$TAG.log | cut -c1-300 | head -n 30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 3 $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$TAG.json"))
    print("fwd ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
    print("gemm frac", d["roofline"]["frac"], "path frac", d["roofline"]["path_frac"], d["roofline"]["class_ms_per_step"], "clocks", d["clocks"])
    t = d["train"]; print("train ms/step", t["ms_per_step"], t["breakdown_ms"], "path_frac", t["path_frac"], t["class_ms_per_step"])
    print({k: (str(v)[:200]) for k, v in t.items() if "drop" in k})
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-3000:])
PY
