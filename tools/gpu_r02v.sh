#!/usr/bin/env bash
# visit r02v (gpurun --gpus N): the driver's multi-GPU launch of the DEFAULT bench line at HEAD (torchrun, one rank per GPU): cfg2 value / e2e,
# train record with the gradient all-reduce, cfg4 sub-record (forward + training step), cfg5 sweep; gradient equality across ranks
set -u
OUT=gpurun_out
TAG=${1:-r02v}
NG=$(nvidia-smi -L | wc -l)
mkdir -p $OUT
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $NG --steps 20 --warmup 5 > $OUT/bench_${NG}gpu_$TAG.json 2> $OUT/bench_${NG}gpu_$TAG.err
echo "bench $NG gpu exit $? wall $(( $(date +%s) - T0 )) s"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${NG}gpu_$TAG.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "e2e_fp32", d["e2e_fp32_io"]["value"], "bank", d["e2e_resident_bank"]["value"], "clk", d["clocks"])
    t = d["train"]; print("train ms", t["ms_per_step"], t["breakdown_ms"], t["allreduce_bytes"])
    c = d["cfg4"]; print("cfg4", c["value"], c["ms_per_step"], "e2e", c["e2e"]["value"], "train", c["train"]["ms_per_step"], c["train"]["breakdown_ms"], c["train"].get("allreduce_bytes"))
    print("sweep", [(r["S"], r["clips_per_gpu"], round(r["ms_per_step"], 1), round(r["path_frac"], 3)) for r in d["sweep_cfg5"]["rows"]])
except Exception as e:
    print("parse failed", e); print(open("$OUT/bench_${NG}gpu_$TAG.err").read()[-3000:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/ddp_grad_check.py --out $OUT/ddp_grad_check_${NG}gpu_$TAG.json > $OUT/ddp_$TAG.log 2>&1
echo "ddp grad check exit $?"; tail -n 2 $OUT/ddp_$TAG.log | cut -c1-1200
