"""Host vs device bag restart loop on the fixture, in one process: where do they part? (GPU box)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import fixture_as_torch, load_golden
import mpgnn_b200
from mpgnn_b200 import search, search_device as sd
fx = fixture_as_torch("fixture_len3")
g = load_golden("search_bags_len3")
data = mpgnn_b200.Data(x=fx["x"], edge_index=fx["edge_index"], edge_type=fx["edge_type"], labels=fx["labels"].unsqueeze(-1),
                       num_nodes=fx["x"].size(0), source_nodes_mask=[])
dev = torch.device("cuda")
for order in ("device_first", "host_first"):
    sg = sd.SearchGraph(fx["edge_index"], fx["edge_type"], fx["x"].size(0), dev)
    graph = search._graph_of(data, dev)
    x_dev = fx["x"].to(dev)
    lab = fx["labels"].float().to(dev)
    state = sd.step0_state(sg, 0, lab, [], "synthetic")
    sd.create_bags(sg, state)
    _, _, ed, dd = mpgnn_b200.score_relation_parallel(data, 0, [], 2, "synthetic")
    bag_data = search._copy_bag(data); search.create_bags(ed, dd, bag_data)
    def run_dev():
        rec = {}
        out = sd.bag_restart_loop(sg, graph, state, 0, x_dev, 2, search.bag_seed(1, 0), record=rec)
        return out[0], rec
    def run_host():
        rec = {}
        out = search.score_relation_bags_parallel(bag_data, 0, 2, "synthetic", metapath_len=1, record=rec)
        return out[1], rec
    runs = [run_dev, run_host] if order == "device_first" else [run_host, run_dev]
    for fn in runs:
        loss, rec = fn()
        print(order, fn.__name__, "loss", loss, "traj[:3]", rec["traj"][:3], "n", len(rec["traj"]), "lin0", rec["lin"][0],
              "ref traj[:3]", g["m0_r0_loss_traj"][:3])
