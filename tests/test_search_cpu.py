"""CPU tests of the search stage: the oracle against the reference goldens, the host-side
restatements (dictionaries, relation discovery, selection rules, partitions) and the multi-rank
fan-out over a world_size-2 gloo group with stand-in scorers."""
import os
import random
import socket
import sys
import types

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, load_golden, fixture_as_torch
from oracle import search_oracle as so

import mpgnn_b200
from mpgnn_b200 import search


def _data(fx):
    return mpgnn_b200.Data(x=fx["x"], edge_index=fx["edge_index"], edge_type=fx["edge_type"],
                           labels=fx["labels"].unsqueeze(-1), num_nodes=fx["x"].size(0), source_nodes_mask=[])


def test_oracle_scorer_matches_reference_golden(fx3):
    g = load_golden("search_len3")
    ei, et, lab = fx3["edge_index"].numpy(), fx3["edge_type"].numpy(), fx3["labels"].numpy()
    rels = so.connected_relations_step0(ei, et, lab)
    assert rels == g["actual_relations"].tolist()
    losses = []
    for r in rels:
        traj, w, arg, keys = so.score_relation(ei, et, r, lab, fx3["x"].size(0))
        assert keys == g["r%d_dest_keys" % r].tolist()                  # dict order drives the RNG stream
        ref = g["r%d_loss_traj" % r]
        assert np.allclose(traj, ref, rtol=1e-5, atol=1e-9)
        assert arg == g["r%d_argmax_dst" % r].tolist()
        assert np.allclose(w[torch.tensor(keys)].numpy(), g["r%d_final_w" % r], atol=1e-6)
        losses.append(traj[-1])
    assert so.gap_select_step0(rels, losses) == g["step0_best"].tolist()


def test_host_dictionaries_and_relations_match_oracle(fx3):
    g = load_golden("search_len3")
    data = _data(fx3)
    rels = search.node_types_and_connected_relations(data, BAGS=False, dataset="synthetic")
    assert rels == g["actual_relations"].tolist()
    for r in rels:
        src = np.unique(fx3["edge_index"].numpy()[0][fx3["edge_type"].numpy() == r]).tolist()
        ed, dd = search.create_edge_dictionary(data, r, src, BAGS=False, dataset="synthetic")
        ed_o, dd_o = so.relation_dictionaries(fx3["edge_index"].numpy(), fx3["edge_type"].numpy(), r,
                                              fx3["labels"].numpy())
        assert list(ed.keys()) == g["r%d_sources" % r].tolist() and ed == ed_o
        assert list(dd.keys()) == g["r%d_dest_keys" % r].tolist() and dd == dd_o
        random.seed(search.SCORER_SEED_BASE + r)
        w = search.initialize_weights(data, dd, BAGS=False)
        assert np.allclose(w[torch.tensor(list(dd.keys()))].numpy(), g["r%d_init_w" % r], atol=0)
    # bag-mode host functions against the oracle restatement
    ei, et, lab = fx3["edge_index"].numpy(), fx3["edge_type"].numpy(), fx3["labels"].numpy()
    ed, dd = so.relation_dictionaries(ei, et, 0, lab)
    bags_o, labels_o = so.create_bags(ed, dd)
    d2 = _data(fx3)
    search.create_bags(ed, dd, d2)
    assert d2.bags == bags_o and d2.bag_labels.reshape(-1).tolist() == [float(v) for v in labels_o]
    assert search.node_types_and_connected_relations(d2, BAGS=True, dataset="synthetic") == \
        so.connected_relations_bags(ei, et, bags_o)
    mask, ed_b, dd_b = so.bag_dictionaries(ei, et, 1, bags_o, labels_o)
    assert mask == search._bag_sources(d2.bags)
    e_p, d_p = search.create_edge_dictionary(d2, 1, mask, BAGS=True, dataset="synthetic")
    assert e_p == ed_b and d_p == dd_b and list(d_p) == list(dd_b) and list(e_p) == list(ed_b)
    kb, kl = so.clean_bags_for_relation_type(bags_o, labels_o, ed_b)
    kb2, kl2 = search.clean_bags_for_relation_type(d2, e_p)
    assert kb == kb2 and kl2.reshape(-1).tolist() == kl
    assert search.accept_bag_relations([(1, 0.1), (2, 0.2)]) == [1, 2]
    assert search.accept_bag_relations([(1, 0.1), (2, 0.2), (3, 0.9)]) == []          # exactly two gaps: nothing
    assert search.accept_bag_relations([(1, 0.1), (2, 0.2), (3, 0.9), (4, 0.95)]) == [1]   # strict `<`


def test_selection_rules_and_partitions():
    # gap rule (main.py:1346-1355): `<=` the value below the largest gap; < 2 gaps keeps everything
    assert search.gap_select_step0([5, 6, 7, 8], [0.30, 0.01, 0.02, 0.31]) == [6, 7]
    assert search.gap_select_step0([1, 2], [0.5, 0.1]) == [1, 2]
    assert search.gap_select_step0([3], [0.2]) == [3]
    assert search.gap_select_step0([1, 2, 3], [0.1, 0.1, 0.9]) == [1, 2]
    # np.array_split / block partition semantics of the reference
    for n in (0, 1, 5, 7, 16):
        for size in (1, 2, 3, 8):
            items = list(range(10, 10 + n))
            parts = [search.relation_split(items, size, r) for r in range(size)]
            assert sum(parts, []) == items
            assert parts == [[int(v) for v in a] for a in np.array_split(np.asarray(items), size)]
            blocks = [search.candidate_block(n, size, r) for r in range(size)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(size - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    # final selection (main.py:1463-1476): stable top-3, greedy union while test F1 strictly improves
    calls = []

    def union(metas):
        calls.append([list(m) for m in metas])
        return {1: 0.8, 2: 0.9, 3: 0.9}[len(metas)]

    fm, f1 = search.final_selection({"[1]": 0.7, "[2, 0]": 0.9, "[3]": 0.9, "[4]": 0.1}, union)
    assert fm == [[2, 0], [3]] and f1 == 0.9 and calls[0] == [[2, 0]] and len(calls) == 3


def _stub_score(data, rel):
    return {0: 0.0, 1: 0.0196, 2: 0.21, 3: 0.22}.get(int(rel), 0.5)


def _stub_eval(meta):
    return 0.5 + 0.1 * meta[0]


def _stub_union(metas):
    return 0.6 + 0.05 * len(metas) if len(metas) < 2 else 0.6


def _worker(rank, size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(size))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=size)
    fx = fixture_as_torch("fixture_len3")
    fx["edge_type"] = fx["edge_type"].clone()
    fx["edge_type"][::7] = 3          # make more relations leave positive nodes
    fx["edge_type"][::11] = 2
    res = search.greedy_search(_data(fx), None, 2, 64, 4, 64, 2, "synthetic", comm=search.Comm(), score_fn=_stub_score,
                               eval_fn=_stub_eval, union_fn=_stub_union)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_fanout_world_size_2_gloo_matches_single_process():
    fx = fixture_as_torch("fixture_len3")
    fx["edge_type"] = fx["edge_type"].clone()
    fx["edge_type"][::7] = 3
    fx["edge_type"][::11] = 2
    single = search.greedy_search(_data(fx), None, 2, 64, 4, 64, 2, "synthetic", comm=search.Comm(),
                                  score_fn=_stub_score, eval_fn=_stub_eval, union_fn=_stub_union)
    assert len(single["relations"]) == 4 and single["kept"] == [r for r in single["relations"] if r in (0, 1)]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):           # every rank derives the same decisions as the single-process run
        assert results[r] == single


def test_oracle_bag_scorer_matches_reference_golden(fx3):
    """Bag mode (restarts, freezing, LinearLayerAttri training) of the oracle against the trajectory
    recorded from the unmodified reference (tests/golden/make_golden_bags.py)."""
    g = load_golden("search_bags_len3")
    ei, et, lab = fx3["edge_index"].numpy(), fx3["edge_type"].numpy(), fx3["labels"].numpy()

    def unr(f, p):
        return [f[p[i]:p[i + 1]].tolist() for i in range(len(p) - 1)]

    ed, dd = so.relation_dictionaries(ei, et, 0, lab)
    bags, labels = so.create_bags(ed, dd)
    assert bags == unr(g["m0_bags_flat"], g["m0_bags_ptr"]) and labels == g["m0_bag_labels"].tolist()
    assert so.connected_relations_bags(ei, et, bags) == g["m0_relations"].tolist()
    rec = {}
    loss, preds, skip, lin = so.score_relation_bags(ei, et, 1, fx3["x"], fx3["x"].size(0), bags, labels,
                                                    so.bag_seed(1, 1), record=rec)
    tag = "m0_r1_"
    ref = g[tag + "loss_traj"]
    assert len(rec["traj"]) == len(ref) and np.allclose(rec["traj"], ref, rtol=1e-4, atol=1e-7)
    assert loss == pytest.approx(float(g[tag + "loss"]), abs=1e-7) and bool(skip) == bool(g[tag + "skip"])
    assert rec["dest_keys"] == g[tag + "dest_keys"].tolist()
    assert rec["bags"] == unr(g[tag + "bags_flat"], g[tag + "bags_ptr"])
    assert rec["frozen_hist"] == unr(g[tag + "frozen_flat"], g[tag + "frozen_ptr"])
    assert np.allclose(np.array(rec["lin_hist"]), g[tag + "lin_hist"], atol=1e-6)
    assert list(preds.keys()) == g[tag + "pred_keys"].tolist()
    vals = np.array([x for k in preds for x in preds[k]])
    assert np.allclose(vals, g[tag + "pred_vals"], atol=1e-5)


def test_create_bags_is_linear_and_equal_to_the_literal_restatement():
    """`create_bags` keeps a set of the bags appended so far instead of the reference's `[value] not in bag` list scan
    (main.py:558, quadratic in the number of bags): same bags, same order, same labels on random dictionaries, and a
    200k-source graph in seconds (the literal loop needs minutes there)."""
    import time
    rng = np.random.default_rng(5)
    for trial in range(20):
        n_src, n_dst = int(rng.integers(1, 60)), int(rng.integers(1, 40))
        ed, dd = {}, {}
        for s in rng.permutation(200)[:n_src].tolist():
            dsts = rng.integers(0, n_dst, size=int(rng.integers(1, 6))).tolist()      # duplicates on purpose
            ed[int(s)] = [int(d) for d in dsts]
            lab = float(rng.integers(0, 2))
            for d in dsts:
                dd.setdefault(int(d), []).append(lab)
        bags_o, labels_o = so.create_bags(ed, dd)
        d2 = types.SimpleNamespace()
        search.create_bags(ed, dd, d2)
        assert d2.bags == bags_o and d2.bag_labels.reshape(-1).tolist() == [float(v) for v in labels_o], trial
    n = 200_000
    src = np.arange(n)
    dst = rng.integers(0, n, size=(n, 3))
    lab = rng.integers(0, 2, size=n).astype(float)
    ed = {int(s): [int(v) for v in dst[s]] for s in src}
    dd = {}
    for s in src:
        for v in dst[s]:
            dd.setdefault(int(v), []).append(float(lab[s]))
    big = types.SimpleNamespace()
    t0 = time.time()
    search.create_bags(ed, dd, big)
    assert time.time() - t0 < 20.0
    assert len(big.bags) == len(big.bag_labels) > n // 4


def test_clean_dictionaries_matches_per_source_dot():
    """Batched feature . weight product with the reference's own torch.dot on the (practically never hit) band around
    the 0.01 threshold: the same sources are dropped as by the per-source loop of main.py:456-477."""
    gen = torch.Generator().manual_seed(3)
    n = 500
    x = torch.nn.functional.one_hot(torch.randint(0, 2, (n,), generator=gen), 2).float()
    x[7] = torch.tensor([0.01, 0.0])                       # exactly on the threshold band
    lin = torch.tensor([[1.0, 0.004]])
    mod = types.SimpleNamespace(output=types.SimpleNamespace(LinearLayerAttri=types.SimpleNamespace(weight=lin)))
    data = types.SimpleNamespace(x=x)
    ed = {int(s): [int(s + 1) % n, int(s + 2) % n] for s in range(0, n, 3)}
    dd = {}
    for s, ds in ed.items():
        for d in ds:
            dd.setdefault(d, []).append(0 if s % 2 else 1)
    dd_in = {k: list(v) for k, v in dd.items()}
    e2, d2 = search.clean_dictionaries(data, ed, dd_in, mod)
    exp_e, exp_d = dict(ed), {k: list(v) for k, v in dd.items()}
    for key in ed:
        if torch.dot(x[key], lin[0]).item() < 0.01:
            for destination in exp_e[key]:
                if 0 in exp_d[destination]:
                    exp_d[destination].remove(0)
            del exp_e[key]
    assert e2 == exp_e and d2 == exp_d and list(e2) == list(exp_e)
    assert 0 < len(e2) < len(ed)


def test_weight_initialisation_equals_the_per_destination_loop():
    """initialize_weights / reinitialize_weights write all destinations with one scatter; the draws (Python `random`,
    dictionary order) and the double -> float32 rounding are those of the reference's per-destination assignment
    (main.py:479-516)."""
    rng = np.random.default_rng(0)
    n = 5000
    dd = {int(k): [float(v) for v in rng.integers(0, 2, size=int(rng.integers(1, 4)))] for k in rng.permutation(n)[:1500]}
    data = types.SimpleNamespace(num_nodes=n)
    random.seed(5)
    got = search.initialize_weights(data, dd, False)
    random.seed(5)
    exp = torch.zeros(n)
    for key, values in dd.items():
        exp[key] = abs(min(values) + random.uniform(-0.2, 0.2))
    assert torch.equal(got, exp)
    frozen = list(dd)[::3]
    random.seed(6)
    got2 = search.reinitialize_weights(data, dd, got.reshape(-1, 1), frozen)
    random.seed(6)
    exp2, fz = torch.zeros(n), set(frozen)
    for key in dd:
        exp2[key] = got[key] if key in fz else random.uniform(0.0, 1.0)
    assert torch.equal(got2, exp2)
    assert torch.equal(search.initialize_weights(data, {}, False), torch.zeros(n))
    assert torch.equal(search.reinitialize_weights(data, {}, got, []), torch.zeros(n))


def _record_scorer(monkeypatch):
    """Replace the device scorer by a recorder: what reaches the GPU is the whole contract of the host side."""
    calls = []

    def fake_run_scorer(graph, relation, weights, node_labels, source_mask=None, epochs=None, lr=None):
        calls.append((int(relation), weights.clone(), node_labels.clone(), None if source_mask is None else source_mask.clone()))
        # deterministic, relation dependent; relations >= 2 sit above a gap the step-0 rule cuts at
        loss = float(weights.double().sum() % 1.0) * 0.001 + 0.01 * int(relation) + (0.5 if int(relation) >= 2 else 0.0)
        return torch.tensor([loss]), weights, None

    monkeypatch.setattr(search, "run_scorer", fake_run_scorer)
    monkeypatch.setattr(search, "_graph_of", lambda data, device: None)
    return calls


@pytest.mark.parametrize("dataset", ["synthetic", "fb15k-237"])
def test_dictionary_free_scoring_hands_the_device_the_same_inputs(fx3, monkeypatch, dataset):
    """score_relation_parallel(dictionaries=False) -- the form greedy_search uses for every relation -- must reach the
    device scorer with bit-identical initial weights, labels and source mask as the dictionary-building form."""
    calls = _record_scorer(monkeypatch)
    data = _data(fx3)
    n = fx3["x"].size(0)
    if dataset == "synthetic":
        masks = [[], np.unique(fx3["edge_index"].numpy()[0])[::2].tolist()]
    else:                                            # labels aligned with the labelled-source list (main.py:424)
        labelled = torch.randperm(n, generator=torch.Generator().manual_seed(1))[: n // 3].tolist()
        data.labels = torch.randint(0, 2, (len(labelled), 1), generator=torch.Generator().manual_seed(2)).float()
        masks = [labelled]
    for mask in masks:
        for rel in range(fx3["num_relations"]):
            del calls[:]
            full = search.score_relation_parallel(data, rel, list(mask), 2, dataset, device="cpu")
            fast = search.score_relation_parallel(data, rel, list(mask), 2, dataset, device="cpu", dictionaries=False)
            assert full[:2] == fast[:2] and fast[2] is None and fast[3] is None and isinstance(full[2], dict)
            (r1, w1, l1, m1), (r2, w2, l2, m2) = calls
            assert r1 == r2 == rel and torch.equal(w1, w2) and torch.equal(l1, l2)
            assert (m1 is None and m2 is None) or torch.equal(m1, m2)
            assert float(w1.abs().sum()) > 0
    # the scatter-min form on its own, against the dictionary form, with repeated destinations and integer labels
    rng = np.random.default_rng(0)
    cols = rng.integers(0, 50, size=400)
    lab = rng.integers(0, 2, size=400)
    dd = {}
    for d, l in zip(cols.tolist(), lab.tolist()):
        dd.setdefault(d, []).append(l)
    random.seed(9)
    exp = search.initialize_weights(types.SimpleNamespace(num_nodes=60), dd, False)
    random.seed(9)
    assert torch.equal(search.initial_weights_from_edges(60, cols, lab), exp)
    assert torch.equal(search.initial_weights_from_edges(60, cols[:0], lab[:0]), torch.zeros(60))


def test_greedy_search_builds_dictionaries_for_kept_relations_only(fx3, monkeypatch):
    """Default step 0 of greedy_search: every relation is scored without dictionaries, the kept ones get theirs from
    create_edge_dictionary -- the same state, bags and candidates as when every score call returns its dictionaries."""
    _record_scorer(monkeypatch)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)      # the default score_fn names the current GPU
    built = []
    real_ced = search.create_edge_dictionary

    def counting_ced(data, relation, mask, BAGS, dataset):
        built.append(int(relation))
        return real_ced(data, relation, mask, BAGS, dataset)

    monkeypatch.setattr(search, "create_edge_dictionary", counting_ced)
    seen_bags = []

    def bag_score_fn(bag_data, rel, mlen):          # record the bags the state produced, then skip the relation
        seen_bags.append((rel, mlen, [list(b) for b in bag_data.bags], bag_data.bag_labels.reshape(-1).tolist()))
        return rel, 1.0, None, {}, True

    def run(score_fn):
        del seen_bags[:]
        data = _data(fx3)
        data.labels = torch.ones_like(data.labels)            # every relation is a candidate of step 0
        res = search.greedy_search(data, None, 2, 64, fx3["num_relations"], 64, 2, "synthetic", score_fn=score_fn,
                                   eval_fn=lambda meta: 0.5 + 0.01 * sum(meta), union_fn=lambda metas: 0.9,
                                   bag_score_fn=bag_score_fn, max_depth=1)
        return res, [tuple(map(str, b)) for b in seen_bags]

    res_fast, bags_fast = run(None)
    kept = res_fast["kept"]
    assert sorted(set(built)) == sorted(kept) and len(kept) < len(res_fast["relations"])    # dictionaries: kept relations only

    def tuple_score_fn(d, rel):
        r = search.score_relation_parallel(d, rel, d.source_nodes_mask, 2, "synthetic", device="cpu")
        return r[1], r[2], r[3]

    res_full, bags_full = run(tuple_score_fn)
    for key in ("relations", "losses", "kept", "candidates", "final_dict", "final_meta"):
        assert res_fast[key] == res_full[key], key
    assert bags_fast == bags_full and len(bags_fast) > 0


def test_lpt_assignment_covers_every_position_once_and_balances():
    metas = [[0], [1], [0], [2, 0], [3, 0], [1, 2, 0], [4, 2, 0], [2, 0], [5, 1, 2, 0]]
    for size in (1, 2, 3, 8):
        parts = search.lpt_assignment(metas, size)
        assert sorted(i for p in parts for i in p) == list(range(len(metas)))
        for p in parts:                                     # all copies of a metapath on one rank
            for i in p:
                assert all(j in p for j, m in enumerate(metas) if m == metas[i])
    two = search.lpt_assignment(metas, 2)
    load = [sum(len(m) for m in {tuple(metas[i]) for i in p}) for p in two]
    assert abs(load[0] - load[1]) <= 1
    assert search.lpt_assignment(metas, 2) == two           # deterministic


def _worker3(rank, size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(size))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import types
    import torch.distributed as dist
    from mpgnn_b200 import search_device as sd
    dist.init_process_group("gloo", rank=rank, world_size=size)
    comm = search.Comm()
    # (a) final selection with a rank per prefix: same result and the same three unions as the sequential rule
    calls = []

    def union(metas):
        calls.append([list(m) for m in metas])
        return {1: 0.8, 2: 0.9, 3: 0.9}[len(metas)]

    fm, f1 = search.final_selection({"[1]": 0.7, "[2, 0]": 0.9, "[3]": 0.9, "[4]": 0.1}, union, comm)
    # (b) the state of an accepted relation travels from the rank that computed it
    pipe = object.__new__(search._DevicePipeline)
    pipe.device, pipe.sd, pipe.sg = torch.device("cpu"), sd, types.SimpleNamespace(n=50)
    owner = 2
    st = None
    if rank == owner:
        st = sd.BagState(7, torch.arange(3, 20, 2), torch.arange(50, dtype=torch.int32) % 3,
                         torch.arange(50, dtype=torch.int32) % 5, (torch.arange(50) % 2).float())
    got = pipe.share_state(st, owner, comm)
    q.put((rank, fm, f1, calls, got.rel, got.src_order.tolist(), got.count0.tolist(), got.count1.tolist(), got.labels.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_3_gloo_final_selection_and_state_broadcast():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker3, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    results = {r[0]: r[1:] for r in (q.get(timeout=120) for _ in range(3))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(3):
        fm, f1, calls, rel, src, c0, c1, lab = results[r]
        assert fm == [[2, 0], [3]] and f1 == 0.9                       # the sequential rule's answer (test above)
        assert calls == [[[2, 0]], [[2, 0], [3]], [[2, 0], [3], [1]]][r:r + 1]      # rank r trained prefix r, nothing else
        assert rel == 7 and src == list(range(3, 20, 2))
        assert c0 == [i % 3 for i in range(50)] and c1 == [i % 5 for i in range(50)] and lab == [float(i % 2) for i in range(50)]


def test_final_dict_is_shared_across_one_vs_rest_label_sets():
    """main.py:1208 creates `final_dict` once, before the loop over the one-vs-rest label sets (a 3-class label file
    gives three of them, main.py:161-172), every label set's candidates are merged into it (a repeated metapath is
    overwritten by the later label set) and the top-3 / greedy-union selection runs ONCE after the loop
    (main.py:1463-1476) -- what mpgnn_b200.main.main() does with `greedy_search(..., final_dict=shared, select=False)`."""
    fx = fixture_as_torch("fixture_len3")
    three_class = (fx["labels"] + (torch.arange(fx["labels"].numel()) % 3 == 0).long()).clamp(max=2)       # classes 0, 1, 2
    binary_sets = [(three_class == c).long() for c in range(3)]
    shared, unions, per_set = {}, [], []
    for c, lab in enumerate(binary_sets):
        data = mpgnn_b200.Data(x=fx["x"], edge_index=fx["edge_index"], edge_type=fx["edge_type"], labels=lab.unsqueeze(-1),
                               num_nodes=fx["x"].size(0), source_nodes_mask=[])
        res = search.greedy_search(data, None, 2, 64, 4, 64, 2, "synthetic", comm=search.Comm(),
                                   score_fn=lambda d, rel: 0.01 * (int(rel) + 1),            # one gap at most: all kept
                                   eval_fn=lambda meta, c=c: 0.3 + 0.1 * c + 0.01 * meta[0],
                                   union_fn=lambda metas: unions.append(metas) or 0.5, final_dict=shared, select=False)
        assert res["final_meta"] is None and res["final_dict"] is shared
        per_set.append(set(str([r]) for r in res["kept"]))
    assert not unions                                               # nothing selected inside the loop
    assert set(shared) == set().union(*per_set)
    for key in shared:                                              # the LAST label set that produced a key wins
        last = max(c for c in range(3) if key in per_set[c])
        assert shared[key] == pytest.approx(0.3 + 0.1 * last + 0.01 * int(key.strip("[]")))
    f_meta, f1 = search.final_selection(shared, lambda metas: unions.append([list(m) for m in metas]) or 0.5)
    best = sorted(shared.items(), key=lambda kv: kv[1], reverse=True)[0][0]
    assert f_meta == [[int(best.strip("[]"))]] and f1 == 0.5 and len(unions) == 2          # second union did not improve
