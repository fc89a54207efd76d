#!/usr/bin/env bash
# Round-2 second visit: tests again, the full bench line (train / cfg4 / sweep / eager sub-records), reference arm, ncu of the training step.
set -u
OUT=gpurun_out
TAG=${1:-r02b}
mkdir -p $OUT
rm -f $OUT/grad_parity.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_real_reference_gpu.py -q -rf -s --no-header -p no:cacheprovider > $OUT/pytest_train_$TAG.log 2>&1
echo "train + real-reference tests exit $?"; grep -E "passed|failed" $OUT/pytest_train_$TAG.log | tail -n 3; grep -E "^FAILED|^\{'" $OUT/pytest_train_$TAG.log | cut -c1-600 | head -n 30
timeout 600 python -m pytest tests/test_gpu_parity.py -q --no-header -p no:cacheprovider > $OUT/pytest_fwd_$TAG.log 2>&1
echo "forward tests exit $?"; tail -n 4 $OUT/pytest_fwd_$TAG.log | cut -c1-300
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$TAG.json"))
    print("fwd ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity"])
    print("gemm frac", d["roofline"]["frac"], "path frac", d["roofline"]["path_frac"], d["roofline"]["class_ms_per_step"])
    t = d["train"]; print("train ms/step", t["ms_per_step"], t["breakdown_ms"], "path_frac", t["path_frac"], t["gemm_tflops"], t["class_ms_per_step"], "tape GB", t["tape_gb"])
    c = d["cfg4"]; print("cfg4 fwd ms", c["ms_per_step"], "value", c["value"], "e2e", c["e2e"]["value"], "train ms", c["train"]["ms_per_step"], c["train"]["class_ms_per_step"])
    print("sweep", [(r["S"], r["clips_per_gpu"], round(r["ms_per_step"], 1), round(r["path_frac"], 3)) for r in d["sweep_cfg5"]["rows"]])
    print("eager", d["gpu_eager_baseline"])
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-3000:])
PY
timeout 400 python bench.py --impl reference --steps 6 --warmup 2 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
cut -c1-700 $OUT/bench_ref_$TAG.json
# ncu: launch list of one training step (cold-cache, serialised: shares only), then full captures of the new kernels
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --kernel-name-base demangled -s 2600 -c 600 --csv --log-file $OUT/launches_train_$TAG.csv \
    python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_list_train_$TAG.log 2>&1
echo "ncu train list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:wgrad_umma2_kernel" -s 30 -c 4 \
    -o $OUT/prof_wgrad_$TAG -f python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_wgrad_$TAG.log 2>&1
echo "ncu wgrad exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attn_bwd" -s 12 -c 2 \
    -o $OUT/prof_attnbwd_$TAG -f python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_attnbwd_$TAG.log 2>&1
echo "ncu attn bwd exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:ln_bwd_kernel|act_bwd_kernel|gelu_fwd_kernel|colsum_kernel" -s 40 -c 6 \
    -o $OUT/prof_trainrows_$TAG -f python bench.py --train-only --steps 3 --warmup 3 > $OUT/ncu_trainrows_$TAG.log 2>&1
echo "ncu train rows exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:assemble_kernel|time_l1_kernel|layernorm_reg_kernel|row_stats_finalize" -s 12 -c 8 -o $OUT/prof_rows_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_rows_$TAG.log 2>&1
echo "ncu rows exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:cast_kernel" -s 420 -c 4 -o $OUT/prof_cast_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_cast_$TAG.log 2>&1
echo "ncu cast exit $?"
for rep in $OUT/prof_wgrad_$TAG $OUT/prof_attnbwd_$TAG $OUT/prof_trainrows_$TAG $OUT/prof_rows_$TAG $OUT/prof_cast_$TAG; do
  [ -f "$rep.ncu-rep" ] && python tools/ncu_summary.py $rep.ncu-rep > $rep.csv 2>/dev/null
done
ls -la $OUT | tail -n 30
