#!/bin/bash
# full ncu capture of the E-step class kernels of one tune.py run (first E-step only)
mkdir -p gpurun_out
export PYLDA_CLASSES="${PYLDA_CLASSES:-8x1,4x2,2x4,1x8}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:estep -c ${NCU_COUNT:-5} -o gpurun_out/${NCU_OUT:-prof_v2} -f \
    python scripts/tune.py ${TUNE_DOCS:-100000} > gpurun_out/ncu_tune.log 2>&1
tail -3 gpurun_out/ncu_tune.log
