"""Pins the oracle (oracle/mpgnn_oracle.py) against (1) the reference's own fixtures
(edges.pkl, labels.pkl -- SURVEY.md section 4) and (2) outputs of the unmodified
reference recorded by tests/golden/make_golden.py.  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import load_golden, rel_err
from oracle import mpgnn_oracle as orc

FP32_TOL = 1e-5  # normalised max error (north star: 1e-5 relative for fp32)


def _check_edges_pkl(fx, name):
    g = load_golden(name + "_edges_pkl")
    n, r_tot = fx["x"].size(0), fx["num_relations"]
    assert int(g["n_rel"]) == r_tot
    ptr, col, perm = orc.relation_csr(fx["edge_index"].numpy(), fx["edge_type"].numpy(), n, r_tot)
    for r in range(r_tot):
        lo, hi = ptr[r * n], ptr[(r + 1) * n]
        rows = np.repeat(np.arange(n), np.diff(ptr[r * n:(r + 1) * n + 1]))
        cols = col[lo:hi]
        ids = np.unique(np.concatenate([rows, cols]))  # gtn_files compacts ids per relation (main.py:212-215)
        m = sp.csr_matrix((np.ones(hi - lo, np.float32), (np.searchsorted(ids, rows), np.searchsorted(ids, cols))),
                          shape=(len(ids), len(ids)))
        m.sum_duplicates()
        m.sort_indices()
        assert tuple(g["shape_%d" % r]) == m.shape
        assert np.array_equal(g["indptr_%d" % r], m.indptr)
        assert np.array_equal(g["indices_%d" % r], m.indices)
        assert np.array_equal(g["data_%d" % r], m.data)  # duplicate edges counted twice


def test_csr_matches_reference_edges_pkl(fx3, fx4):
    _check_edges_pkl(fx3, "fixture_len3")
    _check_edges_pkl(fx4, "fixture_len4")


def test_csr_matches_reference_masked_edge_index(fx3):
    g = load_golden("layer_len3")
    n, r_tot = fx3["x"].size(0), fx3["num_relations"]
    ei = fx3["edge_index"].numpy()
    et = fx3["edge_type"].numpy()
    ptr, col, perm = orc.relation_csr(ei, et, n, r_tot)
    ptr_t, row_t, perm_t = orc.relation_csr(ei, et, n, r_tot, transpose=True)
    for r in range(r_tot):
        mei = g["mei_r%d" % r]  # reference masked_edge_index(edge_index, edge_type == r)
        assert np.array_equal(mei, orc.masked_edge_index(ei, et == r))
        # stable bucketing by row == stable argsort of the filtered list by row
        order = np.argsort(mei[0], kind="stable")
        lo, hi = ptr[r * n], ptr[(r + 1) * n]
        assert np.array_equal(col[lo:hi], mei[1][order])
        assert np.array_equal(np.repeat(np.arange(n), np.diff(ptr[r * n:(r + 1) * n + 1])), mei[0][order])
        order_t = np.argsort(mei[1], kind="stable")
        lo, hi = ptr_t[r * n], ptr_t[(r + 1) * n]
        assert np.array_equal(row_t[lo:hi], mei[0][order_t])


def test_split_matches_reference_labels_pkl():
    for name in ("fixture_len3", "fixture_len4"):
        g = load_golden(name)
        assert np.array_equal(g["labels_pkl_train"][:, 0], g["train_idx"])
        assert np.array_equal(g["labels_pkl_train"][:, 1], g["train_y"])
        assert np.array_equal(g["labels_pkl_val"][:, 0], g["val_idx"])
        assert np.array_equal(g["labels_pkl_test"][:, 1], g["test_y"])
        assert (len(g["train_idx"]), len(g["val_idx"]), len(g["test_idx"])) == (3600, 900, 500)


@pytest.mark.parametrize("tag", ["l0", "l1"])
def test_layer_forward_backward_matches_reference(fx3, tag):
    g = load_golden("layer_len3")
    s = int(g["row_stride"])
    x = fx3["x"] if tag == "l0" else torch.from_numpy(g["x64"])
    g_out = torch.from_numpy(g["g64"])
    w, root, b = (torch.from_numpy(g[tag + "_" + k]) for k in ("weight", "root", "bias"))
    for r in range(fx3["num_relations"]):
        out, h, cnt = orc.conv_forward(x, fx3["edge_index"], fx3["edge_type"], r, w, root, b)
        gx, gw, groot, gb = orc.conv_backward(x, fx3["edge_index"], fx3["edge_type"], r, w, root, h, cnt, g_out)
        pre = "%s_r%d_" % (tag, r)
        assert rel_err(out[::s], g[pre + "out"]) < FP32_TOL
        assert rel_err(gx[::s], g[pre + "gx"]) < FP32_TOL
        assert rel_err(gw, g[pre + "gw"]) < FP32_TOL
        assert rel_err(groot, g[pre + "groot"]) < FP32_TOL
        assert rel_err(gb, g[pre + "gbias"]) < FP32_TOL


def test_model_init_matches_reference_state_dict():
    g = load_golden("model_len3")
    torch.manual_seed(30)
    sd = orc.mpnetm_init(2, 64, 2, [[1, 0]])
    keys = [k[4:] for k in g if k.startswith("sd0.")]
    assert sorted(keys) == sorted(sd.keys())
    for k in keys:
        assert np.array_equal(sd[k].numpy(), g["sd0." + k]), k  # same RNG order => bit-exact


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def test_model_eval_and_train_step_match_reference(fx3):
    g = load_golden("model_len3")
    sd = _sd(g, "sd0.")
    meta = [[1, 0]]
    logp = orc.mpnetm_forward(sd, fx3["x"], fx3["edge_index"], fx3["edge_type"], meta)
    assert rel_err(logp, g["eval_logp"]) < FP32_TOL
    n = fx3["x"].size(0)
    masks = {(0, k): torch.from_numpy(np.unpackbits(g["step_mask_%d" % k], axis=1)[:, :64].astype(np.float32))
             for k in range(2)}
    assert masks[(0, 0)].shape == (n, 64)
    loss, grads, _ = orc.mpnetm_loss_and_grads(sd, fx3["x"], fx3["edge_index"], fx3["edge_type"], meta,
                                               fx3["train_idx"], fx3["train_y"], masks)
    assert abs(float(loss) - float(g["step_loss"])) < FP32_TOL * abs(float(g["step_loss"]))
    for k in sd:
        assert rel_err(grads[k], g["step_grad." + k]) < 5 * FP32_TOL, k
    sd1 = orc.adam_step({k: v.clone() for k, v in sd.items()}, grads, {})
    for k in sd:
        assert rel_err(sd1[k], g["sd1." + k]) < FP32_TOL, k
    logp1 = orc.mpnetm_forward(sd1, fx3["x"], fx3["edge_index"], fx3["edge_type"], meta)
    pred = logp1.argmax(1)
    f1_tr = orc.macro_f1(pred[torch.as_tensor(fx3["train_idx"])], fx3["train_y"])
    f1_va = orc.macro_f1(pred[torch.as_tensor(fx3["val_idx"])], fx3["val_y"])
    assert abs(f1_tr - g["step_val"][0]) < 1e-12 and abs(f1_va - g["step_val"][1]) < 1e-12


def test_dropout_free_training_trace_matches_reference(fx3):
    g = load_golden("model_len3")
    sd = _sd(g, "sd0.")
    f1, sd20, trace = orc.score_candidate(sd, fx3, [[1, 0]], epochs=20, return_trace=True)
    ref = g["trace20_m10"]
    tr = np.array(trace)
    assert np.allclose(tr[:, 0], ref[:, 0], rtol=1e-4), (tr[:, 0], ref[:, 0])  # train loss
    assert np.allclose(tr[:, 1], ref[:, 1], rtol=1e-4)  # val loss
    assert np.allclose(tr[:, 2:], ref[:, 2:], atol=2e-3)  # macro-F1 (a flipped node moves it by ~1e-3)
    assert f1 == pytest.approx(ref[-1, 3], abs=2e-3)


def test_two_metapath_trace_matches_reference(fx3):
    g = load_golden("model_len3")
    torch.manual_seed(30)
    sd = orc.mpnetm_init(2, 64, 2, [[1, 0], [3]])
    _, _, trace = orc.score_candidate(sd, fx3, [[1, 0], [3]], epochs=5, return_trace=True)
    ref = g["trace5_m10_m3"]
    tr = np.array(trace)
    assert np.allclose(tr[:, :2], ref[:, :2], rtol=1e-4)
    assert np.allclose(tr[:, 2:], ref[:, 2:], atol=2e-3)


def test_macro_f1_matches_sklearn():
    from sklearn.metrics import f1_score
    rng = np.random.RandomState(0)
    for c in (2, 3, 5):
        for _ in range(5):
            p, t = rng.randint(0, c, 200), rng.randint(0, c, 200)
            assert orc.macro_f1(p, t) == pytest.approx(f1_score(p, t, average="macro"), abs=1e-12)
    assert orc.macro_f1(np.zeros(10, int), np.zeros(10, int)) == 1.0
    assert orc.macro_f1(np.zeros(10, int), np.ones(10, int)) == 0.0


def test_rgcn_baseline_oracle_matches_reference_golden():
    """f4: the oracle's restatement of `Net` / PyG's RGCNConv against the golden recorded from the unmodified
    model.py behind the stand-ins (tests/golden/make_golden_rgcn.py): same draws, same log-probabilities."""
    import torch
    from conftest import fixture_as_torch, load_golden
    from oracle import mpgnn_oracle as orc
    g = load_golden("rgcn_len3")
    fx = fixture_as_torch("fixture_len3")
    torch.manual_seed(30)
    sd = orc.net_init(2, 64, fx["num_relations"], 64, 2)
    for k, v in sd.items():
        assert torch.equal(v, torch.from_numpy(g["sd0." + k])), k
    for length, key in ((2, "eval_logp"), (3, "eval_logp_len3")):
        logp = orc.net_forward(sd, fx["x"], fx["edge_index"], fx["edge_type"], length)
        ref = torch.from_numpy(g[key])
        assert float((logp - ref).abs().max()) <= 1e-6 * float(ref.abs().max()), key
