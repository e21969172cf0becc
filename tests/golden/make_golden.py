"""Generates tests/golden/*.npz by running the UNMODIFIED reference sources (behind
oracle/ref_shims.py) in the build container.  Run from the repo root:

    python tests/golden/make_golden.py [--skip-long]

Seeds: torch.manual_seed(30) before every model construction (reference main.py:31-32),
torch.manual_seed(0) for random inputs / cotangents, torch.manual_seed(7) before the
captured-dropout train step.  The GPU box has no /root/reference, so tests only read
the committed .npz files.
"""
import argparse
import os
import pickle
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

FIX3 = os.path.join(ref_shims.REFERENCE_ROOT, "data/synthetic/metapath_length_3/overlap_0rels_0/")
FIX4 = os.path.join(ref_shims.REFERENCE_ROOT, "data/synthetic/metapath_length_4/overlap_0_rels_0/")
ROW_STRIDE = 4  # floating-point layer goldens keep every 4th row to stay small


def load_fixture(ref_main, folder):
    labels, features, links, _bl, n_rel = ref_main.load_files(folder + "node.dat", folder + "link.dat",
                                                              folder + "label.dat")
    x = ref_main.get_node_features(features)
    ei, et = ref_main.get_edge_index_and_type_no_reverse(links)
    node_idx, tr_i, tr_y, te_i, te_y, va_i, va_y = ref_main.splitting_node_and_labels(labels, features, [],
                                                                                   "synthetic")
    return dict(x=x, edge_index=ei, edge_type=et, labels=labels, num_relations=n_rel,
                train_idx=tr_i, train_y=tr_y, val_idx=va_i, val_y=va_y, test_idx=te_i, test_y=te_y)


def canonical_edges_pkl(folder):
    with open(folder + "edges.pkl", "rb") as f:
        mats = pickle.load(f)
    out = {"n_rel": np.int64(len(mats))}
    for r, m in enumerate(mats):
        m = m.tocsr()
        m.sum_duplicates()
        m.sort_indices()
        out["indptr_%d" % r] = m.indptr.astype(np.int32)
        out["indices_%d" % r] = m.indices.astype(np.int32)
        out["data_%d" % r] = m.data.astype(np.float32)
        out["shape_%d" % r] = np.array(m.shape, dtype=np.int64)
    return out


def save_fixture(name, fx, folder):
    with open(folder + "labels.pkl", "rb") as f:
        lab_pkl = pickle.load(f)
    # the reference's own fixture pins the split
    assert [p[0] for p in lab_pkl[0]] == [int(v) for v in fx["train_idx"]]
    assert [p[0] for p in lab_pkl[1]] == [int(v) for v in fx["val_idx"]]
    assert [p[0] for p in lab_pkl[2]] == [int(v) for v in fx["test_idx"]]
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        x=fx["x"].numpy(), edge_index=fx["edge_index"].numpy().astype(np.int32),
        edge_type=fx["edge_type"].numpy().astype(np.int32), labels=fx["labels"].numpy().astype(np.int32),
        num_relations=np.int64(fx["num_relations"]),
        train_idx=np.array(fx["train_idx"], dtype=np.int32), train_y=np.array(fx["train_y"], dtype=np.int32),
        val_idx=np.array(fx["val_idx"], dtype=np.int32), val_y=np.array(fx["val_y"], dtype=np.int32),
        test_idx=np.array(fx["test_idx"], dtype=np.int32), test_y=np.array(fx["test_y"], dtype=np.int32),
        labels_pkl_train=np.array(lab_pkl[0], dtype=np.int32), labels_pkl_val=np.array(lab_pkl[1], dtype=np.int32),
        labels_pkl_test=np.array(lab_pkl[2], dtype=np.int32))
    np.savez_compressed(os.path.join(HERE, name + "_edges_pkl.npz"), **canonical_edges_pkl(folder))


def layer_goldens(ref_layer, fx):
    out = {}
    ei, et = fx["edge_index"], fx["edge_type"]
    n = fx["x"].size(0)
    torch.manual_seed(0)
    x64 = torch.randn(n, 64)
    g64 = torch.randn(n, 64)
    out["x64"], out["g64"] = x64.numpy(), g64.numpy()
    for tag, xin, fin in (("l0", fx["x"], 2), ("l1", x64, 64)):
        torch.manual_seed(30)
        conv = ref_layer.CustomRGCNConv(fin, 64, 1, flow="target_to_source")
        out[tag + "_weight"] = conv.weight.detach().numpy().copy()
        out[tag + "_root"] = conv.root.detach().numpy().copy()
        out[tag + "_bias"] = conv.bias.detach().numpy().copy()
        for r in range(fx["num_relations"]):
            xi = xin.clone().requires_grad_(True)
            conv.zero_grad()
            o = conv(0, r, xi, ei, et)
            (o * g64).sum().backward()
            pre = "%s_r%d_" % (tag, r)
            out[pre + "out"] = o.detach().numpy()[::ROW_STRIDE].copy()
            out[pre + "gx"] = xi.grad.numpy()[::ROW_STRIDE].copy()
            out[pre + "gw"] = conv.weight.grad.numpy().copy()
            out[pre + "groot"] = conv.root.grad.numpy().copy()
            out[pre + "gbias"] = conv.bias.grad.numpy().copy()
            tmp = ref_layer.masked_edge_index(ei, et == r)
            out["mei_r%d" % r] = tmp.numpy().astype(np.int32)
    out["row_stride"] = np.int64(ROW_STRIDE)
    np.savez_compressed(os.path.join(HERE, "layer_len3.npz"), **out)


def _data_bag(fx):
    d = types.SimpleNamespace()
    d.x, d.edge_index, d.edge_type = fx["x"], fx["edge_index"], fx["edge_type"]
    d.train_idx, d.val_idx, d.test_idx = fx["train_idx"], fx["val_idx"], fx["test_idx"]
    d.train_y, d.val_y, d.test_y = fx["train_y"], fx["val_y"], fx["test_y"]
    d.num_nodes = fx["x"].size(0)
    return d


def model_goldens(ref_main, ref_model, fx, long_runs):
    out = {}
    data = _data_bag(fx)
    meta = [[1, 0]]
    torch.manual_seed(30)
    net = ref_model.MPNetm(2, 64, fx["num_relations"], 64, 2, 1, meta)
    for k, v in net.state_dict().items():
        out["sd0." + k] = v.numpy().copy()
    net.eval()
    with torch.no_grad():
        out["eval_logp"] = net(data.x, data.edge_index, data.edge_type).numpy().copy()
    # one train step with the dropout masks captured through forward hooks
    masks = []
    hooks = [m.register_forward_hook(lambda mod, i, o: masks.append((o != 0).numpy().copy()))
             for m in (net.dropout, net.dropout2)]
    opt = torch.optim.Adam(net.parameters(), lr=0.01, weight_decay=0.0005)
    torch.manual_seed(7)
    loss, _ = ref_main.mpgnn_train(net, opt, data)
    for h in hooks:
        h.remove()
    out["step_loss"] = np.float32(loss)
    out["step_mask_0"] = np.packbits(masks[0], axis=1)
    out["step_mask_1"] = np.packbits(masks[1], axis=1)
    for k, p in net.named_parameters():
        out["step_grad." + k] = p.grad.numpy().copy()
    for k, v in net.state_dict().items():
        out["sd1." + k] = v.numpy().copy()
    f1_tr, f1_va, _f, loss_val = ref_main.mpgnn_validation(net, data, None)
    out["step_val"] = np.array([f1_tr, f1_va, float(loss_val)], dtype=np.float64)

    # dropout-free traces (nn.Dropout.p set to 0 at run time: configuration, not a source edit)
    def trace(metapaths, epochs):
        torch.manual_seed(30)
        n2 = ref_model.MPNetm(2, 64, fx["num_relations"], 64, 2, len(metapaths), metapaths)
        n2.dropout.p = 0.0
        n2.dropout2.p = 0.0
        o2 = torch.optim.Adam(n2.parameters(), lr=0.01, weight_decay=0.0005)
        tr = []
        for _ in range(epochs):
            l, _ = ref_main.mpgnn_train(n2, o2, data)
            f1t, f1v, _f, lv = ref_main.mpgnn_validation(n2, data, None)
            tr.append((l, float(lv), f1t, f1v))
        lt, f1test = ref_main.mpgnn_test(n2, data, None)
        return np.array(tr, dtype=np.float64), np.array([float(lt), f1test]), n2

    tr, te, n2 = trace(meta, 20)
    out["trace20_m10"], out["trace20_m10_test"] = tr, te
    for k, v in n2.state_dict().items():
        out["sd20." + k] = v.numpy().copy()
    tr, te, _ = trace([[1, 0], [3]], 5)
    out["trace5_m10_m3"], out["trace5_m10_m3_test"] = tr, te
    if long_runs:
        for name, mp in (("m10", [[1, 0]]), ("m0", [[0]]), ("m23", [[2, 3]])):
            tr, te, _ = trace(mp, 999)
            out["trace999_" + name], out["trace999_%s_test" % name] = tr, te
            print("999 epochs", mp, "last val f1", tr[-1, 3], "test", te, flush=True)
    np.savez_compressed(os.path.join(HERE, "model_len3.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-long", action="store_true")
    args = ap.parse_args()
    ref_main, ref_model, ref_layer = ref_shims.import_reference()
    torch.set_num_threads(1)
    fx3 = load_fixture(ref_main, FIX3)
    with open(FIX3 + "node_features.pkl", "rb") as f:
        assert np.array_equal(pickle.load(f), fx3["x"].numpy())
    save_fixture("fixture_len3", fx3, FIX3)
    fx4 = load_fixture(ref_main, FIX4)
    save_fixture("fixture_len4", fx4, FIX4)
    layer_goldens(ref_layer, fx3)
    model_goldens(ref_main, ref_model, fx3, not args.skip_long)
    print("goldens written to", HERE)


if __name__ == "__main__":
    main()
