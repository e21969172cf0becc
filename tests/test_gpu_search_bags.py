"""GPU parity of the bag iterations of the search (K5 bag mode) against the reference goldens
(tests/golden/search_bags_len3.npz) and the behavioural anchor of SURVEY §8c: on the fixture the
greedy search must recover the ground-truth metapath [1, 0]."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import search_oracle as so

import mpgnn_b200
from mpgnn_b200 import search

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1500, method="thread")]


def _data(fx):
    return mpgnn_b200.Data(x=fx["x"], edge_index=fx["edge_index"], edge_type=fx["edge_type"],
                           labels=fx["labels"].unsqueeze(-1), num_nodes=fx["x"].size(0), source_nodes_mask=[])


def _unragged(flat, ptr):
    return [flat[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]


@pytest.mark.parametrize("rel0", [0, 1])
def test_bag_scorer_matches_reference_golden(fx3, rel0):
    g = load_golden("search_bags_len3")
    data = _data(fx3)
    _, _, ed, dd = mpgnn_b200.score_relation_parallel(data, rel0, [], 2, "synthetic")
    bag_data = search._copy_bag(data)
    search.create_bags(ed, dd, bag_data)
    pre = "m%d_" % rel0
    assert bag_data.bags == _unragged(g[pre + "bags_flat"], g[pre + "bags_ptr"])          # bit-exact bag construction
    assert bag_data.bag_labels.reshape(-1).tolist() == g[pre + "bag_labels"].tolist()
    rels = search.node_types_and_connected_relations(bag_data, BAGS=True, dataset="synthetic")
    assert rels == g[pre + "relations"].tolist()
    for rr in rels:
        tag = pre + "r%d_" % rr
        rec = {}
        rel, loss, model, preds, skip = search.score_relation_bags_parallel(bag_data, rr, 2, "synthetic", metapath_len=1,
                                                                            record=rec)
        assert rec["bags"] == _unragged(g[tag + "bags_flat"], g[tag + "bags_ptr"])
        assert rec["dest_keys"] == g[tag + "dest_keys"].tolist()
        assert bool(skip) == bool(g[tag + "skip"])
        ref = g[tag + "loss_traj"]
        got = np.array(rec["traj"])
        # EVERY restart, not only the first: same number of restarts, the whole loss trajectory at the 1e-4 bar (the
        # absolute floor covers losses that train to ~1e-9, where a relative error says nothing), the trained weights
        # after each restart, the Linear weight, and the freeze sets -- exactly, they steer the next restart's draws
        n_restarts = int(g[tag + "n_restarts"])
        assert len(got) == len(ref) == 50 * n_restarts
        assert np.allclose(got, ref, rtol=1e-4, atol=1e-7), float(np.abs(got - ref).max())
        fz_ptr, fz_flat = g[tag + "frozen_ptr"], g[tag + "frozen_flat"]
        keys = g[tag + "dest_keys"]
        for k in range(n_restarts):
            assert rec["frozen"][k] == fz_flat[fz_ptr[k]:fz_ptr[k + 1]].tolist(), (tag, k)
            assert np.allclose(rec["w"][k][keys], g[tag + "w_hist"][k], atol=2e-6), (tag, k)
            assert np.allclose(rec["lin"][k], g[tag + "lin_hist"][2 * k + 1], atol=2e-6), (tag, k)
        assert abs(loss - float(g[tag + "loss"])) <= 1e-4 * float(g[tag + "loss"]) + 1e-8
        assert list(preds.keys()) == g[tag + "pred_keys"].tolist()
        pp, pv = g[tag + "pred_ptr"], g[tag + "pred_vals"]
        for i, key in enumerate(preds):                                  # per-source predictions of every restart
            assert np.allclose(preds[key], pv[pp[i]:pp[i + 1]], atol=1e-5), (tag, key)


def test_greedy_search_recovers_ground_truth_metapath(fx3):
    data = _data(fx3)
    bag = mpgnn_b200.Data(**{k: fx3[k] for k in ("x", "edge_index", "edge_type", "train_idx", "train_y", "val_idx",
                                                  "val_y", "test_idx", "test_y")}, num_nodes=fx3["x"].size(0))
    from mpgnn_b200 import main as m
    ev = lambda meta: (torch.manual_seed(30), m.mpgnn_parallel_multiple(bag, 2, 64, 4, 64, 2, [meta], epochs=300))[1]  # noqa: E731
    un = lambda metas: (torch.manual_seed(30), m.mpgnn_parallel_multiple_x(bag, 2, 64, 4, 64, 2, metas, True,  # noqa: E731
                                                                           epochs=300))[1]
    logs = []
    res = search.greedy_search(data, bag, 2, 64, 4, 64, 2, "synthetic", eval_fn=ev, union_fn=un, log=logs.append,
                               max_depth=1)
    print("\n".join(logs))
    assert res["kept"] == [0, 1]
    assert [1, 0] in res["candidates"]                       # the generator's ground truth (metapath.dat: "1 0")
    assert res["final_dict"]["[1, 0]"] > 0.99
    assert res["final_meta"][0] == [1, 0] and res["test_f1"] > 0.99


@pytest.mark.parametrize("rel0", [0, 1])
def test_device_pipeline_bag_scorer_matches_reference_golden(fx3, rel0):
    """The array form of the search state (search_device.py) feeding K5 from device tensors, against the same reference
    golden as the dictionary form above: bags, candidate relations, and for every relation every restart's loss
    trajectory, freeze set and trained weights."""
    from mpgnn_b200 import search_device as sd
    g = load_golden("search_bags_len3")
    data = _data(fx3)
    dev = torch.device("cuda")
    sg = sd.SearchGraph(fx3["edge_index"], fx3["edge_type"], fx3["x"].size(0), dev)
    graph = search._graph_of(data, dev)
    x_dev = fx3["x"].to(dev).contiguous()
    lab = fx3["labels"].float().to(dev)
    state = sd.step0_state(sg, rel0, lab, [], "synthetic")
    sd.create_bags(sg, state)
    pre = "m%d_" % rel0
    assert state.bags_as_lists() == _unragged(g[pre + "bags_flat"], g[pre + "bags_ptr"])
    assert state.bag_labels.tolist() == g[pre + "bag_labels"].tolist()
    rels = sg.connected_relations(sd.bag_member_mask(sg, state))
    assert rels == g[pre + "relations"].tolist()
    for rr in rels:
        tag = pre + "r%d_" % rr
        rec = {}
        loss, lin, vals, visited, skip = sd.bag_restart_loop(sg, graph, state, rr, x_dev, 2, search.bag_seed(1, rr), record=rec)
        assert rec["dest_keys"].tolist() == g[tag + "dest_keys"].tolist()
        ref = g[tag + "loss_traj"]
        got = np.array(rec["traj"])
        n_restarts = int(g[tag + "n_restarts"])
        assert len(got) == len(ref) == 50 * n_restarts
        assert np.allclose(got, ref, rtol=1e-4, atol=1e-7), float(np.abs(got - ref).max())
        fz_ptr, fz_flat = g[tag + "frozen_ptr"], g[tag + "frozen_flat"]
        keys = g[tag + "dest_keys"]
        for k in range(n_restarts):
            assert sorted(torch.nonzero(rec["frozen"][k]).reshape(-1).tolist()) == sorted(fz_flat[fz_ptr[k]:fz_ptr[k + 1]].tolist())
            assert np.allclose(rec["w"][k].cpu().numpy()[keys], g[tag + "w_hist"][k], atol=2e-6), (tag, k)
        assert abs(loss - float(g[tag + "loss"])) <= 1e-4 * float(g[tag + "loss"]) + 1e-8
        assert bool(skip) == bool(g[tag + "skip"])
        assert visited.tolist() == g[tag + "pred_keys"].tolist()
        pp, pv = g[tag + "pred_ptr"], g[tag + "pred_vals"]
        vals_h = vals.cpu().numpy()
        for i, key in enumerate(visited.tolist()):
            assert np.allclose(vals_h[:, key], pv[pp[i]:pp[i + 1]], atol=1e-5), (tag, key)


@pytest.mark.parametrize("case", ["fixture", "fb"])
def test_device_pipeline_decides_exactly_like_host_pipeline(fx3, case):
    """greedy_search through both pipelines (two bag iterations, candidate evaluation stubbed out): same relations,
    bit-identical losses, same accepted relations and candidate list."""
    if case == "fixture":
        data, f, ds, nrel = _data(fx3), 2, "synthetic", 4
    else:
        g = load_golden("search_fb_small")
        x = torch.from_numpy(g["x"])
        data = mpgnn_b200.Data(x=x, edge_index=torch.from_numpy(g["edge_index"]), edge_type=torch.from_numpy(g["edge_type"]),
                               labels=torch.from_numpy(g["labels"]).unsqueeze(-1), num_nodes=x.size(0),
                               source_nodes_mask=g["labelled"].tolist())
        f, ds, nrel = x.size(1), "fb15k-237", 12
    out = {}
    for pipe in ("host", "device"):
        tm = {}
        out[pipe] = search.greedy_search(data, None, f, 64, nrel, 64, 2, ds, eval_fn=lambda meta: 0.5 + 0.001 * sum(meta),
                                         union_fn=lambda metas: 0.9, max_depth=2 if case == "fixture" else 1, pipeline=pipe,
                                         timings=tm)
        print(case, pipe, "search %.2f s for %d relation scorings" % (tm["search_s"], tm["relations_scored"]))
    h, d = out["host"], out["device"]
    assert h["relations"] == d["relations"] and h["kept"] == d["kept"] and h["losses"] == d["losses"]
    assert h["bag_steps"] == d["bag_steps"]
    assert h["candidates"] == d["candidates"] and h["final_meta"] == d["final_meta"]
