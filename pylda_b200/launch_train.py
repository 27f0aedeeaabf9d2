#!/usr/bin/env python
"""Python-3 mirror of the reference training driver (/root/reference/launch_train.py) for the one
inference mode this package implements: variational Bayes (--inference_mode=2).

    python -m pylda_b200.launch_train --input_directory=<dir with train.dat, voc.dat> \
        --output_directory=<dir> --number_of_topics=K --training_iterations=I --inference_mode=2

Same flags, defaults, option.txt, output directory naming, snapshot files and final pickle as the
reference (launch_train.py:10-67, :118-162, :196-204).  Modes 0 (hybrid) and 1 (Monte Carlo) are
other algorithms and out of scope: they are reported like the reference's
"unrecognized inference mode" branch (:190-192).
"""
import datetime
import optparse
import os
import pickle
import sys

FLAGS = (  # name, type, default, help   (launch_train.py:12-62)
    ("input_directory", "string", None, "input directory [None]"),
    ("output_directory", "string", None, "output directory [None]"),
    ("number_of_topics", "int", -1, "total number of topics [-1]"),
    ("training_iterations", "int", -1, "total number of iterations [-1]"),
    ("snapshot_interval", "int", 10, "snapshot interval [10]"),
    ("alpha_alpha", "float", -1, "hyper-parameter for Dirichlet distribution of topics [1.0/number_of_topics]"),
    ("alpha_beta", "float", -1, "hyper-parameter for Dirichlet distribution of vocabulary [1.0/number_of_types]"),
    ("inference_mode", "int", 0, "inference mode [0 (default): hybrid, 1: monte carlo, 2: variational bayes]"),
    # not in the reference: directory of the on-disk CSR cache of the parsed corpus (pylda_b200/corpus_cache.py);
    # in a multi-process run rank 0 parses once and every rank reads only its own shard of documents
    ("csr_cache", "string", None, "directory of the on-disk CSR cache of the parsed corpus [None: parse in every process]"),
)


def parse_args(argv=None):
    parser = optparse.OptionParser()
    for name, kind, default, text in FLAGS:
        parser.add_option("--" + name, type=kind, dest=name, default=default, help=text)
    options, _ = parser.parse_args(argv)
    return options


def load_documents(path):
    """launch_train.py:102-107 -- one document per line, stripped and lower-cased."""
    with open(path, "r") as stream:
        return [line.strip().lower() for line in stream]


def load_vocabulary(path):
    """launch_train.py:110-116 -- first field of every line, de-duplicated through set()."""
    with open(path, "r") as stream:
        return list(set(line.strip().lower().split()[0] for line in stream))


def main(argv=None):
    options = parse_args(argv)
    assert options.number_of_topics > 0
    assert options.training_iterations > 0
    assert options.snapshot_interval > 0
    assert options.input_directory is not None
    assert options.output_directory is not None
    number_of_topics = options.number_of_topics
    if options.csr_cache:
        os.environ["PYLDA_CSR_CACHE"] = options.csr_cache
    training_iterations = options.training_iterations
    snapshot_interval = options.snapshot_interval
    inference_mode = options.inference_mode

    # one process per GPU (torchrun): every rank trains, rank 0 owns the output directory
    from . import distributed
    rank, world_size, _ = distributed.world()
    input_directory = options.input_directory.rstrip("/")
    corpus_name = os.path.basename(input_directory)
    output_directory = options.output_directory
    for path in (output_directory, os.path.join(output_directory, corpus_name)):
        if rank == 0 and not os.path.exists(path):
            os.mkdir(path)
    output_directory = os.path.join(output_directory, corpus_name)

    train_docs_path = os.path.join(input_directory, "train.dat")
    train_docs = load_documents(train_docs_path)
    print("successfully load all training docs from %s..." % os.path.abspath(train_docs_path))
    vocabulary_path = os.path.join(input_directory, "voc.dat")
    vocab = load_vocabulary(vocabulary_path)
    print("successfully load all the words from %s..." % os.path.abspath(vocabulary_path))

    alpha_alpha = options.alpha_alpha if options.alpha_alpha > 0 else 1.0 / number_of_topics     # :119-121
    alpha_beta = options.alpha_beta if options.alpha_beta > 0 else 1.0 / len(vocab)              # :122-124

    suffix = datetime.datetime.now().strftime("%y%m%d-%H%M%S")                                    # :127-138
    suffix += "-lda-I%d-S%d-K%d-aa%f-ab%f-im%d/" % (training_iterations, snapshot_interval, number_of_topics,
                                                    alpha_alpha, alpha_beta, inference_mode)
    output_directory = os.path.join(output_directory, suffix)
    if rank == 0:
        os.mkdir(os.path.abspath(output_directory))

    settings = [("input_directory", input_directory), ("corpus_name", corpus_name),
                ("training_iterations", "%d" % training_iterations), ("snapshot_interval", str(snapshot_interval)),
                ("number_of_topics", str(number_of_topics)), ("alpha_alpha", str(alpha_alpha)),
                ("alpha_beta", str(alpha_beta)), ("inference_mode", "%d" % inference_mode)]
    if rank == 0:
        with open(output_directory + "option.txt", "w") as f:                                     # :147-162
            for key, value in settings:
                f.write("%s=%s\n" % (key, value))
    rule = "========== ========== ========== ========== =========="
    print(rule)
    print("output_directory=" + output_directory)
    for key, value in settings:
        print("%s=%s" % (key, value))
    print(rule)

    if inference_mode == 2:
        from . import variational_bayes
        lda_inferencer = variational_bayes.VariationalBayes()
    else:
        sys.stderr.write("error: unrecognized inference mode %d... (this package implements mode 2, "
                         "variational Bayes, only)\n" % inference_mode)
        return None

    lda_inferencer._initialize(train_docs, vocab, number_of_topics, alpha_alpha, alpha_beta)       # :194
    for _ in range(training_iterations):                                                           # :196-201
        lda_inferencer.learning()
        if lda_inferencer._counter % snapshot_interval == 0:
            lda_inferencer.materialize()       # multi-process: gathers the gamma shards of all ranks (collective)
            if rank == 0:
                lda_inferencer.export_beta(output_directory + "exp_beta-" + str(lda_inferencer._counter))
                lda_inferencer.export_gamma(output_directory + "exp_gamma-" + str(lda_inferencer._counter))
    lda_inferencer.materialize()
    if rank == 0:
        model_snapshot_path = os.path.join(output_directory, "model-" + str(lda_inferencer._counter))
        with open(model_snapshot_path, "wb") as f:                                                 # :203-204
            pickle.dump(lda_inferencer, f)
    return output_directory


if __name__ == "__main__":
    main()
