#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the bench (both arms), the ncu launch list and the full capture of the
# per-document kernels of one E-step of the bench workload (DRAM traffic).  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONHASHSEED=0
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${TESTS:-1}" = "1" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
fi
if [ "${REF:-1}" = "1" ]; then
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; echo "bench ref rc=$?"
cat gpurun_out/bench_ref.json
fi
if [ "${BENCH:-1}" = "1" ]; then
timeout 1500 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.log; echo "bench rc=$?"
cat gpurun_out/bench.json
fi
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --docs ${NCU_DOCS:-200000} --steps 1 --warmup 1 --no-cpu-baseline --no-warm-lda > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch.log
# DRAM traffic of every per-document kernel launch of one E-step (metrics only: small report) ...
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:estep -c ${NCU_COUNT:-16} --csv --log-file gpurun_out/traffic_estep.csv python scripts/tune.py 1000000 > gpurun_out/ncu_traffic.log 2>&1
# ... and the full capture (with source) of the dominant kernel, the streaming one (first launch of an E-step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:estep_stream -c 1 -o gpurun_out/prof_stream -f \
    python scripts/tune.py 1000000 > gpurun_out/ncu_full.log 2>&1
fi
if [ "${NCU:-1}" = "1" ]; then
# the compact stage for long documents in the warm state (EM iteration 5 of the LDA-drawn corpus): full capture
timeout 900 ncu --set full --clock-control none --import-source on -k regex:estep_longc -s 6 -c 1 -o gpurun_out/prof_longc_warm -f \
    python scripts/tune_lda.py 250000 10 > gpurun_out/ncu_longc_warm.log 2>&1
fi
[ -x scripts/ubench/fp64_lat ] && ./scripts/ubench/fp64_lat > gpurun_out/fp64_ubench.txt 2>&1
tail -3 gpurun_out/pytest_gpu.log
