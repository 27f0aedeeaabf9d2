#!/bin/bash
# usage: tune_sweep.sh "<classes1>" "<classes2>" ...   (runs scripts/tune.py for each PYLDA_CLASSES value)
mkdir -p gpurun_out
: > gpurun_out/tune.log
for c in "$@"; do
  echo "=== PYLDA_CLASSES=$c" >> gpurun_out/tune.log
  PYLDA_CLASSES="$c" timeout 600 python scripts/tune.py ${TUNE_DOCS:-200000} >> gpurun_out/tune.log 2>&1
done
cat gpurun_out/tune.log
