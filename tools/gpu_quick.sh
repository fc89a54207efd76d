#!/usr/bin/env bash
# Short GPU visit while iterating on one kernel: selected tests (no -x, so every failing shape is listed), then a bench line.
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> "<pytest -k expr>"'
set -u
TAG=${1:-q}
KEXPR=${2:-attention}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1
echo "pytest($KEXPR) exit $?"; tail -n 25 $OUT/pytest_$TAG.log
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 600 python bench.py --no-cpu-baseline ${BENCH_ARGS:-} > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
  echo "bench exit $?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$TAG.json"))
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity"])
    print("classes", d["roofline"]["class_ms_per_step"], "gemm TF", d["roofline"]["achieved"])
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-2000:])
PY
fi
