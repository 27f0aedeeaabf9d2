#!/bin/bash
# One gpurun --gpus 8 call: config 3 weak + strong scaling and config 5 (+ contention variant) at 8 GPUs.
set -u
N=${N:-8}
mkdir -p gpurun_out
export PYTHONHASHSEED=0
run() {  # name, bench args
  name=$1; shift
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 \
      bench.py --gpus $N "$@" > gpurun_out/n${N}_$name.json 2> gpurun_out/n${N}_$name.log; echo "$name rc=$?"
  cat gpurun_out/n${N}_$name.json
}
run c3_weak --steps 3 --warmup 3 --no-warm --no-warm-lda --no-cpu-baseline
run c3_strong --steps 3 --warmup 3 --scaling strong --no-warm --no-warm-lda --no-cpu-baseline
run c5_weak --config c5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e
run c5x_weak --config c5x --steps 2 --warmup 3 --no-cpu-baseline --no-e2e
nvidia-smi --query-gpu=index,name,memory.used --format=csv > gpurun_out/n${N}_gpus.txt
