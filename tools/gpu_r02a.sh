#!/usr/bin/env bash
# Round-2 first visit: training-leg parity, the real reference on the GPU, eager baseline, sanitizer, ncu of the HBM-bound kernels.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/smi_r02a.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py -q -rf --no-header -p no:cacheprovider > $OUT/pytest_train_r02a.log 2>&1
echo "train tests exit $?"; tail -n 40 $OUT/pytest_train_r02a.log | cut -c1-300
timeout 900 python -m pytest tests/test_real_reference_gpu.py -q -rf -s --no-header -p no:cacheprovider > $OUT/pytest_realref_r02a.log 2>&1
echo "real-reference tests exit $?"; tail -n 25 $OUT/pytest_realref_r02a.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --no-header -p no:cacheprovider > $OUT/pytest_fwd_r02a.log 2>&1
echo "forward tests exit $?"; tail -n 5 $OUT/pytest_fwd_r02a.log | cut -c1-300
timeout 300 python tools/eager_baseline.py --clips 256 --out $OUT/eager_baseline_r02a.json > $OUT/eager_baseline_r02a.log 2>&1
echo "eager baseline exit $?"; tail -n 1 $OUT/eager_baseline_r02a.log | cut -c1-900
# compute-sanitizer: forward smoke (cfg1) and one tiny training step
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck_smoke_r02a.log 2>&1
echo "memcheck smoke exit $?"; tail -n 4 $OUT/sanitizer_memcheck_smoke_r02a.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/train_smoke.py recog_av_small fp16 fp32 > $OUT/sanitizer_memcheck_train_r02a.log 2>&1
echo "memcheck train exit $?"; tail -n 4 $OUT/sanitizer_memcheck_train_r02a.log | cut -c1-300
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/train_smoke.py recog_av_small fp16 > $OUT/sanitizer_racecheck_train_r02a.log 2>&1
echo "racecheck train exit $?"; tail -n 4 $OUT/sanitizer_racecheck_train_r02a.log | cut -c1-300
# HBM-bound kernels of the forward: full captures (achieved DRAM GB/s against the measured copy bandwidth)
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:assemble_kernel|time_l1_kernel|cast_kernel|layernorm_reg_kernel|row_stats_finalize" -s 30 -c 10 -o $OUT/prof_rows_r02a -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_rows_r02a.log 2>&1
echo "ncu rows exit $?"
[ -f $OUT/prof_rows_r02a.ncu-rep ] && python tools/ncu_summary.py $OUT/prof_rows_r02a.ncu-rep > $OUT/prof_rows_r02a.csv 2>/dev/null
ls -la $OUT | tail -n 20
