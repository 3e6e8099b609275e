#!/bin/bash
# One `ncu --set full` capture + text summary (run on the GPU box through gpurun):
#   bash profiles/capture.sh <tag> <kernel-regex> <skip> <case> <envs> [mode]
set -e
tag=$1; kern=$2; skip=$3; shift 3
python profiles/ncu_case.py "$@" > /dev/null                        # the program must exit 0 without ncu first
ncu --set full --clock-control none --import-source on -k "regex:$kern" -s "$skip" -c 1 -f -o "gpurun_out/$tag" \
    python profiles/ncu_case.py "$@" > "gpurun_out/$tag.ncu.log" 2>&1
python profiles/ncu_summary.py "gpurun_out/$tag.ncu-rep" "$2" > "gpurun_out/$tag.ncu_summary.txt" 2>&1 || true
