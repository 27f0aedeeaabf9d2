#!/bin/bash
# One gpurun call while iterating on kernels: GPU parity tests, then per-class timings of one cold E-step
# under a few environment variants.  Output: gpurun_out/quick_*.log
set -u
mkdir -p gpurun_out
export PYTHONHASHSEED=0
D=${D:-1000000}
if [ "${TESTS:-1}" = "1" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/quick_pytest.log
  tail -15 gpurun_out/quick_pytest.log
fi
: > gpurun_out/quick_tune.log
IFS=';' read -ra VARS <<< "${VARIANTS:-}"
for v in "${VARS[@]:-}"; do
  echo "=== variant: ${v:-default}" >> gpurun_out/quick_tune.log
  env $v python scripts/tune.py $D 2>&1 | grep -E "pylda class|classes=|rror" >> gpurun_out/quick_tune.log
done
cat gpurun_out/quick_tune.log
