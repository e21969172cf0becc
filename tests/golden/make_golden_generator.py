"""Goldens for the synthetic-graph generator, recorded from the UNMODIFIED reference script
`data/synthetic/create_graph_multi_metapath_deterministic.py`.  Run from the repo root:

    python tests/golden/make_golden_generator.py

The script is executed in memory from where it lies under /root/reference with exactly two seams, neither of which
touches its rules:
  * the hard-coded output folder (line 42, '/Users/francescoferrini/...') is redirected to a temporary directory;
  * `random.seed(k)` is called first (the script never seeds its RNG), and `pandas.DataFrame` is wrapped so that the
    edge list handed to the sparsification stage (line 369) is recorded before edges are deleted.
pandas 3 needs `future.infer_string = False` for line 381.
What is kept per case: colours, out-degree draws, the edge list before and after sparsification, the planted
relations / colour order, every stage of the label chain and the labels -- the deterministic stages of the generator
(label chain, sparsification, file formats) are then pinned bit for bit; the random draws are pinned in distribution.
"""
import argparse
import os
import random
import sys
import tempfile
import time

import numpy as np
import pandas as pd

REF = "/root/reference/data/synthetic/create_graph_multi_metapath_deterministic.py"
HERE = os.path.dirname(os.path.abspath(__file__))
HARD_CODED = "'/Users/francescoferrini/VScode/MultirelationalGNN/data/synthetic/metapath_length_3/'"

CASES = [  # name, seed, num_nodes, max_rel_for_node, metapath, overlap, shared_relations
    ("len3_o0r0", 11, 400, 4, "red-blue-red-blue", 0, 0),
    ("len2_o1r2", 12, 300, 3, "blue-blue-red", 1, 2),
    ("len4_o2r1", 13, 350, 5, "red-red-blue-blue-red", 2, 1),
    ("len3_o3r1", 14, 300, 10, "blue-red-red-blue", 3, 1),
    ("len3_o3r3", 15, 250, 6, "red-blue-blue-red", 3, 3),
]


def run_reference(seed, num_nodes, max_rel, metapath, overlap, shared):
    src = open(REF).read()
    assert src.count(HARD_CODED) == 1
    tmp = tempfile.mkdtemp(prefix="mpgnn_gen_")
    src = src.replace(HARD_CODED, repr(tmp + "/"))
    captured = {}
    real_df = pd.DataFrame

    def recording_df(data=None, *a, **k):
        if "pre" not in captured and k.get("columns") == ["source", "relation", "destination"]:
            captured["pre"] = np.array(data, dtype=np.int64).reshape(-1, 3)
        return real_df(data, *a, **k)

    pd.set_option("future.infer_string", False)
    ns = {"__name__": "reference_generator"}
    exec(compile(src, REF, "exec"), ns)
    ns["pd"].DataFrame = recording_df
    args = argparse.Namespace(num_nodes=num_nodes, max_rel_for_node=max_rel, metapath=metapath, overlap=overlap,
                              shared_relations=shared, metapath2=None, metapath3=None)
    random.seed(seed)
    t0 = time.time()
    try:
        ns["main"](args)
    finally:
        pd.DataFrame = real_df
    seconds = time.time() - t0
    folder = os.path.join(tmp, "overlap_%drels_%d" % (overlap, shared))
    files = {name: open(os.path.join(folder, name)).read()
             for name in ("node.dat", "link.dat", "label.dat", "embedding.dat", "metapath.dat")}
    return captured["pre"], files, seconds


def main():
    out = {}
    for name, seed, n, max_rel, metapath, overlap, shared in CASES:
        pre, files, seconds = run_reference(seed, n, max_rel, metapath, overlap, shared)
        node = np.array([[int(v) for v in line.split("\t")] for line in files["node.dat"].splitlines()], dtype=np.int64)
        post = np.array([[int(v) for v in line.split("\t")] for line in files["link.dat"].splitlines()],
                        dtype=np.int64).reshape(-1, 3)
        label = np.array([[int(v) for v in line.split("\t")] for line in files["label.dat"].splitlines()], dtype=np.int64)
        emb = np.array([[int(v) for v in line.split("\t") if v != ""] for line in files["embedding.dat"].splitlines()],
                       dtype=np.int64)
        mp_lines = files["metapath.dat"].split("\n")
        out[name + "_colors"] = node[:, 1:].argmax(axis=1)
        out[name + "_pre"] = pre
        out[name + "_post"] = post
        out[name + "_label"] = label[:, 1]
        out[name + "_embedding"] = emb[:, 1:]                      # stage s of the chain in column s
        out[name + "_meta_reversed"] = np.array([int(v) for v in mp_lines[1].split()], dtype=np.int64)
        out[name + "_colors_reversed"] = np.array([int(v) for v in mp_lines[2].split()], dtype=np.int64)
        out[name + "_args"] = np.array([seed, n, max_rel, overlap, shared], dtype=np.int64)
        out[name + "_metapath"] = np.array(metapath)
        out[name + "_seconds"] = np.array(seconds)
        for fname, text in files.items():
            out[name + "_file_" + fname] = np.array(text)
        print("%-10s nodes %d edges %d -> %d positives %d  reference %.2f s" %
              (name, n, len(pre), len(post), int(label[:, 1].sum()), seconds))
    out["cases"] = np.array([c[0] for c in CASES])
    np.savez_compressed(os.path.join(HERE, "generator.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
