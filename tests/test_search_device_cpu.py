"""The array (device) form of the bag pipeline (mpgnn_b200/search_device.py) against the dictionary functions that
restate the reference one by one (mpgnn_b200/search.py, themselves pinned by the reference goldens): the same torch
code runs here on the CPU with a stand-in for the K5 kernel, so every structure, order and random draw can be compared
exactly -- bags, candidate relations, cleaned bags, destination key order, initial and re-drawn weights, freeze sets,
per-source predictions, relabelling, the cleaned dictionaries of the next level, and the next level's bags."""
import random

import numpy as np
import pytest
import torch

from conftest import load_golden

import mpgnn_b200
from mpgnn_b200 import search, search_device as sd


def test_bulk_draws_are_pythons_own_stream():
    random.seed(2105)
    a = [random.uniform(-0.2, 0.2) for _ in range(7)] + [random.uniform(0.0, 1.0) for _ in range(5)]
    nxt = random.random()
    random.seed(2105)
    b = list(sd.uniform_draws(-0.2, 0.2, 7)) + list(sd.uniform_draws(0.0, 1.0, 5))
    assert a == [float(v) for v in b] and random.random() == nxt           # same doubles, generator left in the same state


def _fake_core(csr, n, bags, labels, x, weights, lin0, grad_mask, use_mask, counter):
    """A deterministic stand-in for one K5 restart (not the real arithmetic): exercises freezing, duplicates, masks."""
    ptr, dst = csr
    w = (weights * 0.5 + 0.3 * grad_mask.to(torch.float32) * (0.5 if use_mask else 1.0)).clamp(0, 1)
    lin = lin0.abs()
    first_dst = dst[ptr[:-1].clamp(max=max(dst.numel() - 1, 0))] if dst.numel() else torch.zeros(n, dtype=torch.int64)
    best_dst = torch.tensor([int(first_dst[b[0]]) for b in bags], dtype=torch.int64)
    diff = torch.tensor([0.0 if (i % 3 == 0 or labels[i] == 1) else 0.5 for i in range(len(bags))])
    src_val = torch.full((n,), float("nan"))
    for b in bags:
        for s in b:
            src_val[s] = w[first_dst[s]] * float((x[s] * lin).sum() + 0.9)
    losses = [0.5, 0.4, 0.45, 0.41, 0.3]
    traj = torch.full((50,), losses[min(counter[0], 4)])
    counter[0] += 1
    return traj, w, lin, best_dst, diff, src_val


def _setup(golden_name):
    if golden_name == "fx3":
        from conftest import fixture_as_torch
        fx = fixture_as_torch("fixture_len3")
        x, ei, et, labels, sources, ds = fx["x"], fx["edge_index"], fx["edge_type"], fx["labels"].float(), [], "synthetic"
    else:
        g = load_golden("search_fb_small")
        x, ei, et = torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"]), torch.from_numpy(g["edge_type"])
        labels, sources, ds = torch.from_numpy(g["labels"]).float(), g["labelled"].tolist(), "fb15k-237"
    data = mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, labels=labels.unsqueeze(-1), num_nodes=x.size(0),
                           source_nodes_mask=sources)
    return data, sd.SearchGraph(ei, et, x.size(0), "cpu"), ds


@pytest.mark.parametrize("name,rel0", [("fx3", 0), ("fx3", 1), ("fb", 0), ("fb", 2)])
def test_array_pipeline_equals_dictionary_pipeline(name, rel0):
    data, sg, ds = _setup(name)
    n, f = sg.n, data.x.size(1)
    lab = data.labels.reshape(-1)
    sources = list(data.source_nodes_mask)
    # ---- step 0: relations, initial weights (same draws), dictionaries
    keep = (lab[sg.rows_all] == 1) if ds == "synthetic" else torch.isin(sg.rows_all, torch.tensor(sources))
    assert sg.connected_relations_from_edge_mask(keep) == search.node_types_and_connected_relations(data, False, ds)
    src_list = sources or np.unique(sg.rel_edges(rel0)[0].numpy()).tolist()
    random.seed(1000 + rel0)
    ed, dd = search.create_edge_dictionary(data, rel0, src_list, BAGS=False, dataset=ds)
    w_dict = search.initialize_weights(data, dd, BAGS=False)
    random.seed(1000 + rel0)
    w_arr, mask, node_labels = sd.step0_inputs(sg, rel0, lab, sources, ds)
    assert torch.equal(w_arr, w_dict)
    state = sd.step0_state(sg, rel0, lab, sources, ds)
    assert state.src_order.tolist() == list(ed)
    for d_, vals in dd.items():
        assert int(state.count0[d_]) == sum(1 for v in vals if v == 0) and int(state.count1[d_]) == sum(1 for v in vals if v != 0)
    # ---- two levels of: create_bags -> relations -> bag scoring of every relation -> accept the first two
    bag_data = search._copy_bag(data)
    for level in range(2):
        search.create_bags(ed, dd, bag_data)
        sd.create_bags(sg, state)
        assert state.bags_as_lists() == bag_data.bags
        assert state.bag_labels.tolist() == bag_data.bag_labels.reshape(-1).tolist()
        rels = search.node_types_and_connected_relations(bag_data, BAGS=True, dataset=ds)
        assert sg.connected_relations(sd.bag_member_mask(sg, state)) == rels
        nxt = None
        for rr in rels[:3]:
            csr = sg.csr(rr)
            c1, c2 = [0], [0]

            def fake_dict(graph, relation, bags, bag_labels, x_dev, weights, lin0, grad_mask, use_mask):
                out = _fake_core(csr, n, bags, bag_labels.reshape(-1).tolist(), data.x, weights, lin0, grad_mask, use_mask, c1)
                return out[0], out[1], out[2], out[3], out[4] ** 2, out[5]

            def fake_arr(graph, relation, prob, x_dev, weights, lin0, grad_mask, use_mask):
                ptr, flat = prob.bag_ptr.tolist(), prob.bag_flat.tolist()
                bags = [flat[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]
                return _fake_core(csr, n, bags, prob.bag_labels.tolist(), data.x, weights, lin0, grad_mask, use_mask, c2)

            rec_d, rec_a = {}, {}
            _, loss_d, model_d, preds_d, skip_d = search.score_relation_bags_parallel(
                bag_data, rr, f, ds, metapath_len=level + 1, record=rec_d, restart_fn=fake_dict)
            loss_a, lin_a, vals_a, visited_a, skip_a = sd.bag_restart_loop(
                sg, None, state, rr, data.x, f, search.bag_seed(level + 1, rr), restart_fn=fake_arr, record=rec_a)
            assert loss_a == loss_d and bool(skip_a) == bool(skip_d)
            assert rec_a["dest_keys"].tolist() == rec_d["dest_keys"]
            prob = rec_a["prob"]
            ptr, flat = prob.bag_ptr.tolist(), prob.bag_flat.tolist()
            assert [flat[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)] == rec_d["bags"]
            assert len(rec_a["w"]) == len(rec_d["w"]) == 4
            for k in range(4):
                assert np.array_equal(rec_a["w"][k].numpy(), rec_d["w"][k])
                assert sorted(torch.nonzero(rec_a["frozen"][k]).reshape(-1).tolist()) == sorted(rec_d["frozen"][k])
            assert visited_a.tolist() == list(preds_d)
            for i, s in enumerate(preds_d):
                assert vals_a[:, s].tolist() == preds_d[s]
            assert torch.equal(lin_a, model_d.output.LinearLayerAttri.weight[0])
            if nxt is None:
                # accept this relation: retrain, relabel, new dictionaries, clean
                c1[0] = c2[0] = 0
                dc = search._copy_bag(bag_data)
                preds = search.retrain_bags(dc, rr, {k: list(v) for k, v in preds_d.items()}, True, f, ds,
                                            metapath_len=level + 1, restart_fn=fake_dict)
                src_mask, _ = search.relabel_nodes_inside_bags(preds, dc, model_d)
                e2, d2 = search.create_edge_dictionary(dc, rr, src_mask, BAGS=False, dataset=ds)
                e3, d3 = search.clean_dictionaries(dc, e2, d2, model_d)
                _, _, vals_r, _, _ = sd.bag_restart_loop(sg, None, state.copy(), rr, data.x, f,
                                                         search.bag_seed(level + 1, rr) + search.RETRAIN_SEED_SHIFT,
                                                         max_restarts=1, restart_fn=fake_arr)
                new = sd.accept_relation(sg, state, rr, lin_a, torch.cat([vals_a, vals_r]), visited_a, data.x, ds)
                assert new.labels.tolist() == dc.labels.reshape(-1).tolist()
                assert new.src_order.tolist() == list(e3)
                for d_, vals in d3.items():
                    assert int(new.count0[d_]) == sum(1 for v in vals if v == 0), d_
                    assert int(new.count1[d_]) == sum(1 for v in vals if v != 0), d_
                nxt = (new, e3, d3, dc)
        if nxt is None:
            break
        state, ed, dd, bag_data = nxt
