#!/bin/bash
# One gpurun --gpus N call: multi-GPU parity tests, then bench.py under torchrun (weak and strong scaling).
# Output: gpurun_out/multi_*.{log,json}
set -u
N=${N:-2}
mkdir -p gpurun_out
export PYTHONHASHSEED=0
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/multi_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi_pytest_n$N.log
tail -5 gpurun_out/multi_pytest_n$N.log
fi
for mode in ${MODES:-weak strong}; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 \
      bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 --scaling $mode ${BENCH_ARGS:---no-warm --no-warm-lda --no-cpu-baseline} \
      > gpurun_out/multi_bench_n${N}_$mode.json 2> gpurun_out/multi_bench_n${N}_$mode.log; echo "bench $mode rc=$?"
  cat gpurun_out/multi_bench_n${N}_$mode.json
done
