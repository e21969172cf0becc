"""Search-stage goldens for the NON-synthetic branch (dataset='fb15k-237': labelled source list, labels aligned with
the position in that list, main.py:77-81, 424, 653-654, 1433) recorded from the UNMODIFIED reference behind
oracle/ref_shims.py on a small FB15K-237-shaped graph (the real link.dat is not in the mount).  Run from the repo root:

    python tests/golden/make_golden_search_fb.py

Recorded: step 0 (candidate relations, per-relation loss, kept relations), then for the first kept relation one full
bag iteration exactly as main.py:1385-1435 does it -- create_bags, candidate relations, bag-mode scores, acceptance,
and for every accepted relation retrain_bags -> relabel_nodes_inside_bags -> create_edge_dictionary(dataset=
'fb15k-237') -> clean_dictionaries.  Seams as in make_golden_search.py / make_golden_bags.py.
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

SCORER_SEED_BASE, BAG_SEED_BASE, RETRAIN_SEED_SHIFT = 1000, 2000, 50
DATASET = "fb15k-237"


def make_graph():
    """1500 entities, 12 relations with skewed frequencies, 8 float features, 400 labelled entities whose label is
    planted two hops away: positive iff some relation-2 neighbour has a relation-5 edge to an entity with x[0] > 0.6."""
    g = torch.Generator().manual_seed(237)
    n, r, e, f, n_lab = 1500, 12, 9000, 8, 400
    ei = torch.randint(0, n, (2, e), generator=g)
    et = (torch.rand(e, generator=g) ** 2 * r).long().clamp_(max=r - 1)
    x = torch.rand(n, f, generator=g)
    hot = x[:, 0] > 0.6
    mid = torch.zeros(n, dtype=torch.bool)
    sel5 = et == 5
    mid[ei[0][sel5][hot[ei[1][sel5]]]] = True
    pos = torch.zeros(n, dtype=torch.bool)
    sel2 = et == 2
    pos[ei[0][sel2][mid[ei[1][sel2]]]] = True
    # labelled entities: every second positive plus random others, in a fixed shuffled order
    cand = torch.randperm(n, generator=g)
    labelled = cand[:n_lab].tolist()
    labels = pos[cand[:n_lab]].long()
    x[cand[:n_lab]] = 0                                     # sn() zeroes the features of labelled nodes (main.py:357-364)
    return x, ei, et, labelled, labels


def ragged(lists):
    flat = np.array([v for l in lists for v in l], dtype=np.float64)
    ptr = np.cumsum([0] + [len(l) for l in lists]).astype(np.int64)
    return flat, ptr


def main():
    m, _, _ = ref_shims.import_reference()
    torch.set_num_threads(1)
    x, ei, et, labelled, labels = make_graph()
    f = x.size(1)
    data = ref_shims._Data()
    data.x, data.edge_index, data.edge_type = x, ei, et
    data.labels = labels.unsqueeze(-1)
    data.num_nodes = x.size(0)
    data.bags, data.bag_labels = torch.empty(1), torch.empty(1)
    data.source_nodes_mask = labelled
    out = {"x": x.numpy(), "edge_index": ei.numpy(), "edge_type": et.numpy(), "labelled": np.array(labelled, dtype=np.int64),
           "labels": labels.numpy()}
    rels = m.node_types_and_connected_relations(data, BAGS=False, dataset=DATASET)
    out["actual_relations"] = np.array(rels, dtype=np.int64)
    results = []
    for rel in rels:
        random.seed(SCORER_SEED_BASE + int(rel))
        torch.manual_seed(SCORER_SEED_BASE + int(rel))
        res = m.score_relation_parallel(data, rel, data.source_nodes_mask, f, dataset=DATASET)
        results.append(res)
        print("step 0 relation", rel, "loss", res[1], flush=True)
    losses = [r_[1] for r_ in results]
    out["step0_losses"] = np.array(losses, dtype=np.float64)
    accs = sorted(losses)
    diffs = np.diff(accs)
    best = [r_ for r_ in results if r_[1] <= accs[int(np.argmax(diffs))]] if len(diffs) >= 2 else list(results)
    out["step0_best"] = np.array([b[0] for b in best], dtype=np.int64)
    print("kept", [b[0] for b in best])
    # ---- one bag iteration for the first kept relation (main.py:1381-1435) ----
    rel0, _, edg, dst = best[0]
    out["rel0"] = np.int64(rel0)
    out["rel0_edge_keys"] = np.array(list(edg.keys()), dtype=np.int64)
    out["rel0_dest_keys"] = np.array(list(dst.keys()), dtype=np.int64)
    out["rel0_dest_vals"], out["rel0_dest_ptr"] = ragged(list(dst.values()))
    bag_data = copy.copy(data)
    m.create_bags(edg, dst, bag_data)
    out["bags_flat"], out["bags_ptr"] = ragged(bag_data.bags)
    out["bag_labels"] = bag_data.bag_labels.squeeze(-1).numpy().copy()
    rels_k = m.node_types_and_connected_relations(bag_data, BAGS=True, dataset=DATASET)
    out["bag_relations"] = np.array(rels_k, dtype=np.int64)
    final_result = []
    for rr in rels_k:
        seed = BAG_SEED_BASE + 100 * 1 + rr
        random.seed(seed)
        torch.manual_seed(seed)
        res = m.score_relation_bags_parallel(bag_data, rr, f, dataset=DATASET)
        print("bag relation", rr, "loss", res[1], "skip", res[4], flush=True)
        out["bag_r%d_loss" % rr] = np.float64(res[1])
        out["bag_r%d_skip" % rr] = np.int64(bool(res[4]))
        if res[4] is not True:
            final_result.append(res)
    arr = sorted(r_[1] for r_ in final_result)
    diffs = np.diff(arr)
    idx = int(np.argmax(diffs)) if len(diffs) > 2 else None
    accepted = []
    for res in final_result:
        if (len(diffs) > 2 and res[1] < arr[idx]) or len(diffs) in (0, 1):
            accepted.append(res[0])
            rr = res[0]
            data_copy = copy.copy(bag_data)
            seed = BAG_SEED_BASE + 100 * 1 + rr + RETRAIN_SEED_SHIFT
            random.seed(seed)
            torch.manual_seed(seed)
            preds = m.retrain_bags(data_copy, rr, res[3], BAGS=True, features_dim=f, dataset=DATASET)
            src_mask, new_labels = m.relabel_nodes_inside_bags(preds, data_copy, res[2])
            e2, d2 = m.create_edge_dictionary(data_copy, rr, src_mask, BAGS=False, dataset=DATASET)      # main.py:1433
            e3, d3 = m.clean_dictionaries(data_copy, e2, d2, res[2])
            tag = "acc_r%d_" % rr
            out[tag + "src_mask"] = np.array(src_mask, dtype=np.int64)
            out[tag + "new_labels"] = new_labels.squeeze(-1).numpy().copy()
            out[tag + "edge_keys"] = np.array(list(e3.keys()), dtype=np.int64)
            out[tag + "edge_vals"], out[tag + "edge_ptr"] = ragged(list(e3.values()))
            out[tag + "dest_keys"] = np.array(list(d3.keys()), dtype=np.int64)
            out[tag + "dest_vals"], out[tag + "dest_ptr"] = ragged(list(d3.values()))
            out[tag + "lin"] = res[2].output.LinearLayerAttri.weight.detach().numpy().reshape(-1).copy()
            print("accepted", rr, "sources", len(e3), "destinations", len(d3), "positives", int(new_labels.sum()), flush=True)
    out["bag_accepted"] = np.array(accepted, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "search_fb_small.npz"), **out)
    print("written")


if __name__ == "__main__":
    main()
