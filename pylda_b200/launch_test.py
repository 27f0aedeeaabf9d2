#!/usr/bin/env python
"""Python-3 mirror of the reference held-out driver (/root/reference/launch_test.py): load
test.dat, un-pickle model snapshots, run VariationalBayes.inference (the E-step's held-out branch,
variational_bayes.py:154-155, :202-204, :216) and write gamma with numpy.savetxt.

    python -m pylda_b200.launch_test --input_directory=<dir with test.dat> \
        --model_directory=<launch_train output dir> [--snapshot_index=N]
"""
import optparse
import os
import pickle
import sys

import numpy


def parse_args(argv=None):
    parser = optparse.OptionParser()       # launch_test.py:10-26
    parser.add_option("--input_directory", type="string", dest="input_directory", default=None,
                      help="input directory [None]")
    parser.add_option("--model_directory", type="string", dest="model_directory", default=None,
                      help="model directory [None]")
    parser.add_option("--snapshot_index", type="int", dest="snapshot_index", default=-1,
                      help="snapshot index [-: evaluate on all available snapshots]")
    options, _ = parser.parse_args(argv)
    return options


def evaluate_snapshot(input_snapshot_path, test_docs, output_gamma_path):
    """launch_test.py:90-97."""
    with open(input_snapshot_path, "rb") as f:
        lda_inferencer = pickle.load(f)
    log_likelihood, gamma_values = lda_inferencer.inference(test_docs)
    print("held-out likelihood of snapshot %s is %g" % (os.path.abspath(input_snapshot_path), log_likelihood))
    numpy.savetxt(output_gamma_path, gamma_values)
    return log_likelihood, gamma_values


def main(argv=None):
    options = parse_args(argv)
    assert options.input_directory is not None
    assert options.model_directory is not None
    input_directory = options.input_directory.rstrip("/")
    input_corpus_name = os.path.basename(input_directory)
    model_directory = options.model_directory.rstrip("/")
    if not os.path.exists(model_directory):
        sys.stderr.write("error: model directory %s does not exist...\n" % os.path.abspath(model_directory))
        return None
    # <output>/<corpus>/<run>/ : the corpus name is the parent directory of the run (launch_test.py:45-49)
    model_corpus_name = os.path.basename(os.path.dirname(os.path.abspath(model_directory)))
    if input_corpus_name != model_corpus_name:
        sys.stderr.write("error: corpus name does not match for input (%s) and model (%s)...\n"
                         % (input_corpus_name, model_corpus_name))
        return None
    snapshot_index = options.snapshot_index
    rule = "========== ========== ========== ========== =========="
    print(rule)
    print("model_directory=" + model_directory)
    print("input_directory=" + input_directory)
    print("corpus_name=" + input_corpus_name)
    print("snapshot_index=" + str(snapshot_index))
    print(rule)

    test_docs_path = os.path.join(input_directory, "test.dat")
    with open(test_docs_path, "r") as stream:
        test_docs = [line.strip().lower() for line in stream]
    print("successfully load all testing docs from %s..." % os.path.abspath(test_docs_path))

    results = {}
    if snapshot_index >= 0:
        input_snapshot_path = os.path.join(model_directory, "model-%d" % snapshot_index)
        if not os.path.exists(input_snapshot_path):
            sys.stderr.write("error: model snapshot %s does not exist...\n" % os.path.abspath(input_snapshot_path))
            return None
        results[snapshot_index] = evaluate_snapshot(input_snapshot_path, test_docs,
                                                    os.path.join(model_directory, "test-%d" % snapshot_index))
    else:
        for model_snapshot in sorted(os.listdir(model_directory)):
            if not model_snapshot.startswith("model-"):
                continue
            index = int(model_snapshot.split("-")[-1])
            results[index] = evaluate_snapshot(os.path.join(model_directory, model_snapshot), test_docs,
                                               os.path.join(model_directory, "test-%d" % index))
    return results


if __name__ == "__main__":
    main()
