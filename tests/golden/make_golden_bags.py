"""Bag-iteration goldens of the search stage (bag mode) recorded from the
UNMODIFIED reference behind oracle/ref_shims.py.  Run from the repo root (takes ~15 minutes):

    python tests/golden/make_golden_bags.py

Seam: `random.seed(s); torch.manual_seed(s)` with s = BAG_SEED_BASE + 100*len(metapath) + relation
right before a relation is scored in bag mode (the reference seeds neither).
The body of score_relation_bags_parallel (main.py:853-917) is replayed here with the reference's own
functions so that per-epoch losses and per-restart state can be recorded.
"""
import copy
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

SCORER_SEED_BASE = 1000
BAG_SEED_BASE = 2000
FIX3 = os.path.join(ref_shims.REFERENCE_ROOT, "data/synthetic/metapath_length_3/overlap_0rels_0/")


def ragged(lists):
    flat = np.array([v for l in lists for v in l], dtype=np.int64)
    ptr = np.cumsum([0] + [len(l) for l in lists]).astype(np.int64)
    return flat, ptr


def score_bags_recorded(m, data_object, relation, features_dim, out, tag):
    """main.py:853-917 line by line, with recording."""
    rest = 0
    source_nodes_mask = []
    for bag in data_object.bags:
        for elm in bag:
            if elm not in source_nodes_mask:
                source_nodes_mask.append(elm)
    current_loss = 100
    edge_dictionary, destination_dictionary = m.create_edge_dictionary(data_object, relation, source_nodes_mask,
                                                                       BAGS=True, dataset="synthetic")
    bags, bag_labels = m.clean_bags_for_relation_type(data_object, edge_dictionary)
    weights = m.initialize_weights(data_object, destination_dictionary, BAGS=True)
    grad_mask = torch.ones(len(weights), 1)
    criterion, criterion_per_node = m.get_loss(), m.get_loss_per_node()
    predictions_for_each_restart = {}
    frozen = []
    v = False
    if len(bags) == 1:
        v = True
    if len(bags) > 1 and bag_labels.squeeze().tolist().count(1) == 0:
        v = True
    out[tag + "dest_keys"] = np.array(list(destination_dictionary.keys()), dtype=np.int64)
    out[tag + "bags_flat"], out[tag + "bags_ptr"] = ragged(bags)
    out[tag + "bag_labels"] = bag_labels.squeeze(-1).numpy().copy()
    keys_t = torch.tensor(list(destination_dictionary.keys()))
    out[tag + "init_w"] = weights[keys_t].numpy().copy()
    traj, lin_hist, frozen_hist, w_hist = [], [], [], []
    n_restart = 0
    while rest < 2:
        model = m.get_model(weights, features_dim)
        lin_hist.append(model.output.LinearLayerAttri.weight.detach().numpy().copy().reshape(-1))
        optimizer = m.get_optimizer(model)
        for epoch in range(50):
            loss, mds, loss_per_bag, mdb, preds = m.train(data_object, edge_dictionary, model, optimizer, criterion,
                                                          source_nodes_mask, criterion_per_node, frozen, weights,
                                                          grad_mask, BAGS=True, bags_to_predict=bags,
                                                          bags_to_predict_labels=bag_labels, dataset="synthetic")
            traj.append(float(loss))
        for key, value in mds.items():
            predictions_for_each_restart.setdefault(key, []).append(value.item())
        if loss.item() < current_loss:
            frozen = m.retrieve_destinations_low_loss(mdb, loss_per_bag, source_nodes_mask)
            current_loss = loss.item()
            rest = 0
        else:
            rest += 1
        for node in frozen:
            grad_mask[node] = 0
        frozen_hist.append(np.array(frozen, dtype=np.int64))
        lin_hist.append(model.output.LinearLayerAttri.weight.detach().numpy().copy().reshape(-1))
        w_hist.append(model.input.weights.detach()[keys_t, 0].numpy().copy())
        weights = m.reinitialize_weights(data_object, destination_dictionary, model.input.weights.detach(), frozen,
                                         BAGS=False)
        n_restart += 1
    out[tag + "loss_traj"] = np.array(traj, dtype=np.float64)
    out[tag + "n_restarts"] = np.int64(n_restart)
    out[tag + "lin_hist"] = np.array(lin_hist, dtype=np.float32)
    out[tag + "w_hist"] = np.array(w_hist, dtype=np.float32)
    fz_flat, fz_ptr = ragged([f.tolist() for f in frozen_hist])
    out[tag + "frozen_flat"], out[tag + "frozen_ptr"] = fz_flat, fz_ptr
    out[tag + "loss"] = np.float64(current_loss)
    out[tag + "skip"] = np.int64(v)
    pk = list(predictions_for_each_restart.keys())
    out[tag + "pred_keys"] = np.array(pk, dtype=np.int64)
    pf, pp = ragged([[0] * len(predictions_for_each_restart[k]) for k in pk])
    out[tag + "pred_ptr"] = pp
    out[tag + "pred_vals"] = np.array([x for k in pk for x in predictions_for_each_restart[k]], dtype=np.float64)
    return relation, current_loss, model, predictions_for_each_restart, v


def main():
    m, _, _ = ref_shims.import_reference()
    torch.set_num_threads(1)
    labels, features, links, bl, n_rel = m.load_files(FIX3 + "node.dat", FIX3 + "link.dat", FIX3 + "label.dat")
    x = m.get_node_features(features)
    ei, et = m.get_edge_index_and_type_no_reverse(links)
    data = ref_shims._Data()
    data.x, data.edge_index, data.edge_type = x, ei, et
    data.labels = bl[0].unsqueeze(-1)
    data.num_nodes = x.size(0)
    data.bags, data.bag_labels = torch.empty(1), torch.empty(1)
    data.source_nodes_mask = []
    out = {}
    for rel0 in (0, 1):
        random.seed(SCORER_SEED_BASE + rel0)
        torch.manual_seed(SCORER_SEED_BASE + rel0)
        r = m.score_relation_parallel(data, rel0, [], 2, "synthetic")
        dc = copy.copy(data)
        m.create_bags(r[2], r[3], dc)                                         # main.py:1385
        pre = "m%d_" % rel0
        out[pre + "bags_flat"], out[pre + "bags_ptr"] = ragged(dc.bags)
        out[pre + "bag_labels"] = dc.bag_labels.squeeze(-1).numpy().copy()
        rels = m.node_types_and_connected_relations(dc, BAGS=True, dataset="synthetic")   # main.py:1386
        out[pre + "relations"] = np.array(rels, dtype=np.int64)
        print("metapath", [rel0], "bags", len(dc.bags), "relations", rels, flush=True)
        for rr in rels:
            seed = BAG_SEED_BASE + 100 * 1 + rr
            random.seed(seed)
            torch.manual_seed(seed)
            res = score_bags_recorded(m, dc, rr, 2, out, pre + "r%d_" % rr)
            print("  relation", rr, "loss", res[1], "skip", res[4], "restarts", int(out[pre + "r%d_n_restarts" % rr]),
                  flush=True)
            np.savez_compressed(os.path.join(HERE, "search_bags_len3.npz"), **out)
    print("done")


if __name__ == "__main__":
    main()
