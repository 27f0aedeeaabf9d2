#!/bin/bash
# Tuning aid: per-class kernel times of one cold E-step at several trip caps (where does the time go:
# once-per-document work, full-width trips, compact trips).  Output: gpurun_out/probe_trips.log
set -u
mkdir -p gpurun_out
export PYTHONHASHSEED=0
D=${1:-1000000}
for it in 1 3 6 10 20 50; do
  echo "=== max_iter=$it" 
  TUNE_MAXITER=$it python scripts/tune.py $D 2>&1 | grep -E "pylda class|classes="
done > gpurun_out/probe_trips.log 2>&1
echo "=== max_iter=50 PYLDA_COMPACT=0" >> gpurun_out/probe_trips.log
PYLDA_COMPACT=0 python scripts/tune.py $D 2>&1 | grep -E "pylda class|classes=" >> gpurun_out/probe_trips.log
cat gpurun_out/probe_trips.log
