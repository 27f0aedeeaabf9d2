#!/bin/bash
# 2-GPU check of the CLI under a real torchrun launch: toy corpus, 3 iterations.
set -e
export PYTHONHASHSEED=0
T=$(mktemp -d)
python - "$T" <<'PY'
import sys, os, numpy
sys.path.insert(0, os.getcwd())
from pylda_b200 import synthetic
T = sys.argv[1]
V = 500
rp, ids, cts = synthetic.synthetic_corpus(400, V, seed=5, length="zipf")
docs = synthetic.render_text(rp, ids, cts)
os.mkdir(os.path.join(T, "toy"))
open(os.path.join(T, "toy", "train.dat"), "w").write("\n".join(docs) + "\n")
open(os.path.join(T, "toy", "voc.dat"), "w").write("".join("w%d\t1\n" % i for i in range(V)))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
    -m pylda_b200.launch_train --input_directory=$T/toy --output_directory=$T/out2 --number_of_topics=20 \
    --training_iterations=3 --snapshot_interval=3 --inference_mode=2 2>&1 | grep -E "log likelihood|rror" | sort | uniq -c
python -m pylda_b200.launch_train --input_directory=$T/toy --output_directory=$T/out1 --number_of_topics=20 \
    --training_iterations=3 --snapshot_interval=3 --inference_mode=2 2>&1 | grep -E "log likelihood|rror"
ls $T/out2/toy/*/
