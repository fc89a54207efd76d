#!/usr/bin/env bash
# One `ncu --set full` capture of a single launch picked by a regex on the demangled kernel name.
#   bash tools/ncu_kernel.sh <out-name> <regex> <skip> -- <command...>
set -u
NAME=$1; REGEX=$2; SKIP=$3; shift 4
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$REGEX" -s "$SKIP" -c 1 \
    -o gpurun_out/$NAME -f "$@" > gpurun_out/$NAME.log 2>&1
echo "ncu($NAME) exit $?"; tail -n 3 gpurun_out/$NAME.log
