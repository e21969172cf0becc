"""Search-stage goldens (step 0 of the greedy search) recorded from the
UNMODIFIED reference behind oracle/ref_shims.py.  Run from the repo root:

    python tests/golden/make_golden_search.py

The reference leaves Python's `random` unseeded (main.py:494); the injection seam used here and by
the product is `random.seed(SCORER_SEED_BASE + relation)` immediately before each relation is
scored, so a relation's result does not depend on which rank/GPU scored it.
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

SCORER_SEED_BASE = 1000
FIX3 = os.path.join(ref_shims.REFERENCE_ROOT, "data/synthetic/metapath_length_3/overlap_0rels_0/")


def main():
    ref_main, ref_model, _ = ref_shims.import_reference()
    torch.set_num_threads(1)
    labels, features, links, binary_labels, n_rel = ref_main.load_files(FIX3 + "node.dat", FIX3 + "link.dat",
                                                                        FIX3 + "label.dat")
    x = ref_main.get_node_features(features)
    ei, et = ref_main.get_edge_index_and_type_no_reverse(links)
    data = ref_shims._Data()
    data.x, data.edge_index, data.edge_type = x, ei, et
    data.labels = binary_labels[0].unsqueeze(-1)            # main.py:1283-1284
    data.num_nodes = x.size(0)
    data.bags, data.bag_labels = torch.empty(1), torch.empty(1)
    data.source_nodes_mask = []
    out = {}
    rels = ref_main.node_types_and_connected_relations(data, BAGS=False, dataset="synthetic")   # main.py:1289
    out["actual_relations"] = np.array(rels, dtype=np.int64)
    losses = []
    for rel in rels:
        random.seed(SCORER_SEED_BASE + int(rel))
        torch.manual_seed(SCORER_SEED_BASE + int(rel))
        # body of score_relation_parallel (main.py:727-760), epoch by epoch so the trajectory is recorded
        src = ref_main.masked_edge_index(ei, et == rel)
        source_nodes = torch.unique(src[0]).tolist()
        edge_dict, dest_dict = ref_main.create_edge_dictionary(data, rel, source_nodes, BAGS=False, dataset="synthetic")
        weights = ref_main.initialize_weights(data, dest_dict, BAGS=False)
        keys = np.array(list(dest_dict.keys()), dtype=np.int64)
        out["r%d_dest_keys" % rel] = keys
        out["r%d_init_w" % rel] = weights[torch.from_numpy(keys)].numpy().copy()
        out["r%d_sources" % rel] = np.array(list(edge_dict.keys()), dtype=np.int64)
        model = ref_main.get_model(weights, x.size(1))
        opt = ref_main.get_optimizer(model)
        crit, crit_n = ref_main.get_loss(), ref_main.get_loss_per_node()
        traj = []
        for _ in range(100):
            loss, mds, lpn, mdb, preds = ref_main.train(data, edge_dict, model, opt, crit, source_nodes, crit_n, [],
                                                        weights, torch.tensor(0), BAGS=False, bags_to_predict=None,
                                                        bags_to_predict_labels=None, dataset="synthetic")
            traj.append(float(loss))
        out["r%d_loss_traj" % rel] = np.array(traj, dtype=np.float64)
        out["r%d_final_w" % rel] = model.input.weights.detach()[torch.from_numpy(keys), 0].numpy().copy()
        out["r%d_argmax_dst" % rel] = np.array([mds[s] for s in edge_dict.keys()], dtype=np.int64)
        losses.append(traj[-1])
        print("relation", rel, "sources", len(edge_dict), "dests", len(keys), "final loss", traj[-1], flush=True)
    # step-0 gap rule exactly as main.py:1348-1355
    accs = sorted(losses)
    diffs = np.diff(accs)
    if len(diffs) >= 2:
        idx = int(np.argmax(diffs))
        best = [r for r, l in zip(rels, losses) if l <= accs[idx]]
    else:
        best = list(rels)
    out["step0_losses"] = np.array(losses, dtype=np.float64)
    out["step0_best"] = np.array(best, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "search_len3.npz"), **out)
    print("actual relations", rels, "losses", losses, "kept", best)


if __name__ == "__main__":
    main()
