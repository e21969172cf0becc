"""Goldens of the reference's all-relation RGCN comparison model (`Net`, model.py:132-151; training calls
main.py:1055-1115, the same functions main_rgcn.py defines) recorded from the UNMODIFIED reference behind
oracle/ref_shims.py.  torch_geometric.nn.RGCNConv is third-party: the stand-in restates PyG 2.3.1's per-relation loop
(see ref_shims._RGCNConv).  Run from the repo root:   python tests/golden/make_golden_rgcn.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_shims  # noqa: E402


def main():
    ref_main, ref_model, _ = ref_shims.import_reference()
    torch.set_num_threads(1)
    g = dict(np.load(os.path.join(HERE, "fixture_len3.npz")))
    d = types.SimpleNamespace()
    d.x = torch.from_numpy(g["x"])
    d.edge_index = torch.from_numpy(g["edge_index"].astype(np.int64))
    d.edge_type = torch.from_numpy(g["edge_type"].astype(np.int64))
    for k in ("train", "val", "test"):
        setattr(d, k + "_idx", g[k + "_idx"].astype(np.int64).tolist())
        setattr(d, k + "_y", torch.from_numpy(g[k + "_y"].astype(np.int64)))
    d.num_nodes = d.x.size(0)
    r = int(g["num_relations"])
    out = {}
    torch.manual_seed(30)
    net = ref_model.Net(2, 64, r, 64, 2, 2)
    for k, v in net.state_dict().items():
        out["sd0." + k] = v.numpy().copy()
    net.eval()
    with torch.no_grad():
        out["eval_logp"] = net(d.x, d.edge_index, d.edge_type).numpy().copy()
    opt = torch.optim.Adam(net.parameters(), lr=0.01, weight_decay=0.0005)
    loss, _ = ref_main.mpgnn_train(net, opt, d)
    out["step_loss"] = np.float32(loss)
    for k, p in net.named_parameters():
        out["step_grad." + k] = p.grad.numpy().copy()
    trace = []
    f1t, f1v, _f, lv = ref_main.mpgnn_validation(net, d, None)
    trace.append((loss, float(lv), f1t, f1v))
    for _ in range(19):
        l, _ = ref_main.mpgnn_train(net, opt, d)
        f1t, f1v, _f, lv = ref_main.mpgnn_validation(net, d, None)
        trace.append((l, float(lv), f1t, f1v))
    lt, f1test = ref_main.mpgnn_test(net, d, None)
    out["trace20"] = np.array(trace, dtype=np.float64)
    out["trace20_test"] = np.array([float(lt), f1test])
    # three-layer variant (conv2 applied twice, model.py:143-147)
    torch.manual_seed(30)
    net3 = ref_model.Net(2, 64, r, 64, 2, 3)
    net3.eval()
    with torch.no_grad():
        out["eval_logp_len3"] = net3(d.x, d.edge_index, d.edge_type).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "rgcn_len3.npz"), **out)
    print("written; last epoch", trace[-1], "test", lt, f1test)


if __name__ == "__main__":
    main()
