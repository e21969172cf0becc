"""How close is the bag-mode scorer (K5) to the reference goldens, restart by restart?  (GPU box)"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden, fixture_as_torch
import mpgnn_b200
from mpgnn_b200 import search

fx = fixture_as_torch("fixture_len3")
g = load_golden("search_bags_len3")
data = mpgnn_b200.Data(x=fx["x"], edge_index=fx["edge_index"], edge_type=fx["edge_type"],
                       labels=fx["labels"].unsqueeze(-1), num_nodes=fx["x"].size(0), source_nodes_mask=[])
for rel0 in (0, 1):
    _, _, ed, dd = mpgnn_b200.score_relation_parallel(data, rel0, [], 2, "synthetic")
    bag_data = search._copy_bag(data)
    search.create_bags(ed, dd, bag_data)
    pre = "m%d_" % rel0
    for rr in search.node_types_and_connected_relations(bag_data, BAGS=True, dataset="synthetic"):
        tag = pre + "r%d_" % rr
        rec = {}
        rel, loss, model, preds, skip = search.score_relation_bags_parallel(bag_data, rr, 2, "synthetic", metapath_len=1, record=rec)
        ref = g[tag + "loss_traj"]
        got = np.array(rec["traj"])
        print(tag, "restarts ours", len(got) // 50, "ref", int(g[tag + "n_restarts"]), "loss ours %.6g ref %.6g" % (loss, float(g[tag + "loss"])))
        keys = g[tag + "dest_keys"]
        fz_ref = [g[tag + "frozen_flat"][g[tag + "frozen_ptr"][i]:g[tag + "frozen_ptr"][i + 1]].tolist() for i in range(len(g[tag + "frozen_ptr"]) - 1)]
        for k in range(min(len(got), len(ref)) // 50):
            a, b = got[50 * k:50 * k + 50], ref[50 * k:50 * k + 50]
            rel_e = np.abs(a - b) / np.maximum(np.abs(b), 1e-12)
            fz = rec["frozen"][k] if k < len(rec.get("frozen", [])) else None
            wd = np.abs(rec["w"][k][keys] - g[tag + "w_hist"][k]).max() if k < len(rec.get("w", [])) else None
            lin = np.abs(rec["lin"][k] - g[tag + "lin_hist"][2 * k + 1]).max()
            print("   restart %d: traj max rel err %.3g (first at epoch %s), last loss ours %.8g ref %.8g; frozen equal %s (|ours| %s |ref| %s, symdiff %s); w max abs diff %s; lin diff %.3g"
                  % (k, rel_e.max(), int(np.argmax(rel_e > 1e-5)) if (rel_e > 1e-5).any() else None, a[-1], b[-1],
                     fz == fz_ref[k] if fz is not None else None, len(fz) if fz is not None else None, len(fz_ref[k]),
                     len(set(fz) ^ set(fz_ref[k])) if fz is not None else None, wd, lin))
