#!/usr/bin/env bash
# visit r02c: re-test the rebuilt backward kernels, bench (train record), launch list; under --gpus 2: gradient equality + 2-GPU bench
set -u
OUT=gpurun_out
TAG=${1:-r02c}
NG=$(nvidia-smi -L | wc -l)
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -q -rf --no-header -p no:cacheprovider -k "attention_bwd or golden or oracle_full or named_configs or dropin or loud" > $OUT/pytest_train_$TAG.log 2>&1
echo "train tests exit $?"; grep -E "passed|failed" $OUT/pytest_train_$TAG.log | tail -n 2; grep -E "^FAILED" $OUT/pytest_train_$TAG.log | cut -c1-300 | head
timeout 600 python -m pytest tests/test_gpu_parity.py -q --no-header -p no:cacheprovider -k "fold or host or dropin" > $OUT/pytest_fwd_$TAG.log 2>&1
echo "forward (host / fold / dropin) tests exit $?"; tail -n 3 $OUT/pytest_fwd_$TAG.log | cut -c1-300; grep -E "^FAILED|Error" $OUT/pytest_fwd_$TAG.log | cut -c1-300 | head
timeout 600 python bench.py --train-only --steps 6 > $OUT/bench_train_$TAG.json 2> $OUT/bench_train_$TAG.err
echo "train bench exit $?"; cut -c1-1800 $OUT/bench_train_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -s 2600 -c 320 --csv \
    --log-file $OUT/launches_train_$TAG.csv python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_list_train_$TAG.log 2>&1
echo "ncu train list exit $?"
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/ddp_grad_check.py --out $OUT/ddp_grad_check_$TAG.json > $OUT/ddp_$TAG.log 2>&1
  echo "ddp grad check exit $?"; tail -n 2 $OUT/ddp_$TAG.log | cut -c1-1200
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $NG --steps 20 --warmup 5 > $OUT/bench_${NG}gpu_$TAG.json 2> $OUT/bench_${NG}gpu_$TAG.err
  echo "bench $NG gpu exit $?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${NG}gpu_$TAG.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "e2e_fp32", d["e2e_fp32_io"]["value"], "e2e16", d["e2e_fp32_logits"]["value"])
    t = d["train"]; print("train ms", t["ms_per_step"], t["breakdown_ms"], t["allreduce_bytes"])
    c = d["cfg4"]; print("cfg4", c["value"], c["ms_per_step"], "e2e", c["e2e"]["value"], "train", c["train"]["ms_per_step"], c["train"]["breakdown_ms"])
except Exception as e:
    print("parse failed", e); print(open("$OUT/bench_${NG}gpu_$TAG.err").read()[-3000:])
PY
else
  timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
  echo "bench exit $?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$TAG.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "e2e_fp32", d["e2e_fp32_io"]["value"], "e2e16", d["e2e_fp32_logits"]["value"])
    t = d["train"]; print("train ms", t["ms_per_step"], t["breakdown_ms"], t["class_ms_per_step"])
    c = d["cfg4"]; print("cfg4", c["value"], c["ms_per_step"], "e2e", c["e2e"]["value"], "train", c["train"]["ms_per_step"])
except Exception as e:
    print("parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-3000:])
PY
fi
