#!/bin/bash
# One gpurun call: ncu --set full capture (with source) of the kernels matching $1 on a D=$2 corpus.
# Output: gpurun_out/prof_$3.ncu-rep (read here with ncu -i ... --page source --csv | scripts/ncu_blocks.py)
set -u
mkdir -p gpurun_out
export PYTHONHASHSEED=0
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${SKIP:-0} -c ${COUNT:-2} -o gpurun_out/prof_$3 -f \
    python scripts/tune.py $2 > gpurun_out/prof_$3.log 2>&1
tail -3 gpurun_out/prof_$3.log
ls -la gpurun_out/
