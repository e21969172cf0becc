"""Native candidate trainer (CUDA-graph replays of the whole epoch) against the reference's
recorded training traces and against the Python/autograd path."""
import time

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

import mpgnn_b200

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]


def _bag(fx):
    return mpgnn_b200.Data(**{k: fx[k] for k in ("x", "edge_index", "edge_type", "train_idx", "train_y", "val_idx",
                                                  "val_y", "test_idx", "test_y")}, num_nodes=fx["x"].size(0))


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def test_trainer_dropout_free_trace_matches_reference_golden(fx3):
    g = load_golden("model_len3")
    data = _bag(fx3)
    ref = g["trace20_m10"]
    traces = {}
    for use_graph in (True, False):
        tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, [1, 0], dropout_p=0.0, max_epochs=20)
        tr.load_state_dict(_sd(g, "sd0."))
        got = tr.state_dict()
        for k, v in _sd(g, "sd0.").items():
            assert torch.equal(got[k], v), k                      # flat layout round trip incl. fc transposes
        trace = tr.run(20, use_graph=use_graph)[:20]
        traces[use_graph] = trace
        assert np.allclose(trace[:, 0], ref[:, 0], rtol=1e-4), (trace[:, 0], ref[:, 0])      # train loss
        assert np.allclose(trace[:, 1], ref[:, 1], rtol=1e-4)                                 # validation loss
        assert np.allclose(trace[:, 2:], ref[:, 2:], atol=2e-3)                               # macro-F1 train / val
        assert tr.last_val_f1 == trace[-1, 3]
        sd20 = tr.state_dict()
        for k, v in _sd(g, "sd20.").items():
            assert rel_err(sd20[k], v) < 1e-3, k
        loss_t, f1_t = tr.evaluate("test")
        assert abs(loss_t - g["trace20_m10_test"][0]) < 1e-3 * abs(g["trace20_m10_test"][0])
        assert abs(f1_t - g["trace20_m10_test"][1]) < 2e-3
    assert np.array_equal(traces[True], traces[False])           # graph replay == eager launches, bit for bit


def test_trainer_dropout_seeded_and_deterministic(fx3):
    g = load_golden("model_len3")
    data = _bag(fx3)

    def run(seed):
        tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, [1, 0], dropout_p=0.6, seed=seed, max_epochs=30)
        tr.load_state_dict(_sd(g, "sd0."))
        return tr.run(30)[:30]

    a, b, c = run(5), run(5), run(6)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a[-1, 0] < a[0, 0] and np.all(np.isfinite(a))


def test_mpgnn_parallel_multiple_native_full_candidate(fx3):
    """One candidate scored exactly as the reference does (999 epochs, dropout 0.6, last-epoch
    validation macro-F1): the ground-truth metapath [1,0] must separate the classes."""
    data = _bag(fx3)
    torch.manual_seed(30)
    torch.cuda.synchronize()
    t0 = time.time()
    f1 = mpgnn_b200.mpgnn_parallel_multiple(data, 2, 64, 4, 64, 2, [[1, 0]])
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("candidate [1,0]: val macro-F1 %.4f in %.3f s (%.2f candidates/s)" % (f1, dt, 1.0 / dt))
    assert f1 > 0.99
    torch.manual_seed(30)
    f1_test = mpgnn_b200.mpgnn_parallel_multiple_x(data, 2, 64, 4, 64, 2, [2, 3], True, epochs=200)
    assert 0.5 < f1_test <= 1.0


def test_candidate_batch_runs_concurrently_and_matches_one_by_one(fx3):
    """mpgnn_parallel_multiple_batch (independent trainers on their own streams, host threads) returns
    exactly the validation F1 of the one-candidate-at-a-time calls under the same seed seam."""
    data = _bag(fx3)
    cands = [[1, 0], [0], [1], [0, 1], [1, 1, 0]]
    one_by_one = []
    for c in cands:
        torch.manual_seed(30)
        one_by_one.append(mpgnn_b200.mpgnn_parallel_multiple(data, 2, 64, 4, 64, 2, [c], epochs=40))
    batch = mpgnn_b200.mpgnn_parallel_multiple_batch(data, 2, 64, 4, 64, 2, cands, epochs=40, seed=30, max_concurrent=4)
    assert batch == one_by_one
    assert len(set(batch)) > 1                      # different metapaths do train to different numbers


def test_trainer_on_a_graph_with_hub_buckets_graph_replay_equals_eager(fx3):
    """Hub buckets (> 256 edges) take the chunked aggregation path, which uses stream-ordered scratch: it has to
    work inside the captured epoch graph exactly as in eager launches."""
    ei, et = fx3["edge_index"].clone(), fx3["edge_type"].clone()
    n = fx3["x"].size(0)
    gen = torch.Generator().manual_seed(9)
    hub_src = torch.randint(0, n, (700,), generator=gen)
    ei = torch.cat([ei, torch.stack([torch.full((700,), 11), hub_src]), torch.stack([hub_src, torch.full((700,), 42)])], 1)
    et = torch.cat([et, torch.full((700,), 1), torch.full((700,), 0)])      # hub target in relation 1, hub source in 0
    data = mpgnn_b200.Data(x=fx3["x"], edge_index=ei, edge_type=et, num_nodes=n,
                           **{k: fx3[k] for k in ("train_idx", "train_y", "val_idx", "val_y", "test_idx", "test_y")})
    traces = {}
    for use_graph in (True, False):
        torch.manual_seed(30)
        model = mpgnn_b200.MPNetm(2, 64, 4, 64, 2, 1, [[1, 0]], device="cpu")
        tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, [1, 0], dropout_p=0.6, seed=5, max_epochs=12)
        tr.load_state_dict(model.state_dict())
        traces[use_graph] = tr.run(12, use_graph=use_graph)[:12]
        assert np.all(np.isfinite(traces[use_graph]))
    assert np.array_equal(traces[True], traces[False])
    assert traces[True][-1, 0] < traces[True][0, 0]


def test_planted_metapath_of_a_generated_graph_scores_highest():
    """End to end over the data boundary: a graph from the seeded generator (the reference's synthetic-data rules,
    `mpgnn_b200.synthetic`) handed over as tensors, three candidates through the fan-out call.  The planted metapath
    separates the classes, a candidate that differs in the first hop partly, an unrelated one not at all (macro-F1 of
    the majority class) -- the ordering the greedy search relies on (CPU oracle, no dropout: 0.997 / 0.874 / 0.473)."""
    from mpgnn_b200 import synthetic
    g = synthetic.generate(6000, 5, "red-blue-red-blue", 0, 1, seed=7)
    x, ei, et, y = g.tensors()
    n = g.num_nodes
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(0))
    n_te, n_va = n // 10, (n - n // 10) // 5
    data = mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, num_nodes=n,
                           test_idx=perm[:n_te], test_y=y[perm[:n_te]], val_idx=perm[n_te:n_te + n_va],
                           val_y=y[perm[n_te:n_te + n_va]], train_idx=perm[n_te + n_va:], train_y=y[perm[n_te + n_va:]])
    p = g.planted_relations
    assert p == [3, 5, 3]                                      # the generator is seeded
    cands = [p, [p[0] ^ 1, p[1], p[2]], [7, 0, 5]]
    f1 = mpgnn_b200.mpgnn_parallel_multiple_batch(data, 2, 64, 8, 64, 2, cands, epochs=150, seed=30)
    assert f1[0] > 0.95, f1
    assert f1[1] < f1[0] - 0.03, f1
    assert f1[2] < 0.6, f1


@pytest.mark.parametrize("name,meta", [("m10", [1, 0]), ("m0", [0]), ("m23", [2, 3])])
def test_trainer_999_epoch_score_matches_reference_golden(fx3, name, meta):
    """What the selection consumes (main.py:1134: the LAST epoch's validation macro-F1 after 999 epochs, and the test
    macro-F1 of main.py:1112) against the unmodified reference's 999-epoch dropout-free runs for the ground-truth
    metapath, a one-hop candidate and an unrelated one (tests/golden/make_golden.py: torch.manual_seed(30), p = 0)."""
    g = load_golden("model_len3")
    data = _bag(fx3)
    ref, ref_test = g["trace999_" + name], g["trace999_%s_test" % name]
    torch.manual_seed(30)
    model = mpgnn_b200.MPNetm(2, 64, fx3["num_relations"], 64, 2, 1, [meta], device="cpu")
    if name == "m10":
        for k, v in _sd(g, "sd0.").items():
            assert torch.equal(model.state_dict()[k], v), k       # same seed, same draw order as the reference
    tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, meta, dropout_p=0.0, max_epochs=999)
    tr.load_state_dict(model.state_dict())
    trace = tr.run(999)
    dev_loss = np.abs(trace[:, 0] - ref[:, 0]) / np.abs(ref[:, 0])
    print("%s: train-loss rel dev max over epochs 1-100 %.2e, 1-999 %.2e; last val F1 ours %.6f ref %.6f"
          % (name, dev_loss[:100].max(), dev_loss.max(), trace[-1, 3], ref[-1, 3]))
    assert np.allclose(trace[:100, 0], ref[:100, 0], rtol=1e-3)              # early trajectory: fp32-class agreement
    assert np.allclose(trace[:100, 1], ref[:100, 1], rtol=1e-3)
    assert np.allclose(trace[:, 0], ref[:, 0], rtol=5e-2, atol=1e-4)         # whole run: same optimisation path
    one_node = 1.5 / min(len(fx3["val_idx"]), len(fx3["test_idx"]))           # macro-F1 moves ~1/n per flipped node
    assert abs(tr.last_val_f1 - ref[-1, 3]) <= 2 * one_node, (tr.last_val_f1, ref[-1, 3])
    assert abs(trace[-1, 2] - ref[-1, 2]) <= 2 * one_node                    # train macro-F1, last epoch
    loss_t, f1_t = tr.evaluate("test")
    assert abs(f1_t - ref_test[1]) <= 2 * one_node, (f1_t, ref_test[1])
    assert abs(loss_t - ref_test[0]) <= 5e-2 * abs(ref_test[0]) + 1e-4


def test_trainer_c2_shape_trace_matches_oracle():
    """BASELINE configs[1] shape (generated graph, 100k nodes, 10 relations, planted length-3 metapath, hidden 64):
    30 dropout-free epochs of the native trainer against the CPU oracle's restatement of mpgnn_train /
    mpgnn_validation (oracle/mpgnn_oracle.py: score_candidate) from the same initial state_dict."""
    from oracle import mpgnn_oracle as orc
    from mpgnn_b200 import synthetic
    sg = synthetic.generate(100_000, 10, "red-blue-red-blue", 0, 2, seed=1)
    x, ei, et, y = sg.tensors()
    n = sg.num_nodes
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1))
    n_te, n_va = n // 10, (n - n // 10) // 5
    idx = {"test": perm[:n_te], "val": perm[n_te:n_te + n_va], "train": perm[n_te + n_va:]}
    data = mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, num_nodes=n,
                           **{k + "_idx": v for k, v in idx.items()}, **{k + "_y": y[v] for k, v in idx.items()})
    meta = sg.planted_relations
    torch.manual_seed(30)
    sd = orc.mpnetm_init(2, 64, 2, [meta])
    epochs = 30
    bag = {"x": x, "edge_index": ei, "edge_type": et, "train_idx": idx["train"].tolist(), "train_y": y[idx["train"]],
           "val_idx": idx["val"].tolist(), "val_y": y[idx["val"]]}
    ref = np.array(orc.score_candidate(sd, bag, [meta], epochs=epochs, return_trace=True)[2])
    tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, meta, dropout_p=0.0, max_epochs=epochs)
    tr.load_state_dict(sd)
    trace = tr.run(epochs)[:epochs]
    print("C2 trace: train-loss rel dev max %.2e, val-loss %.2e, val-F1 abs dev max %.2e"
          % ((np.abs(trace[:, 0] - ref[:, 0]) / np.abs(ref[:, 0])).max(),
             (np.abs(trace[:, 1] - ref[:, 1]) / np.abs(ref[:, 1])).max(), np.abs(trace[:, 3] - ref[:, 3]).max()))
    # the first steps at the fp32 bar; over 30 Adam steps rounding differences compound (measured 1.3e-4 at epoch 30)
    assert np.allclose(trace[:5, :2], ref[:5, :2], rtol=1e-5)
    assert np.allclose(trace[:, :2], ref[:, :2], rtol=5e-4)
    assert np.allclose(trace[:, 2:], ref[:, 2:], atol=5e-4)


def test_trainer_multi_metapath_matches_reference_golden(fx3):
    """mpgnn_parallel_multiple_x (main.py:1136-1160) on the native trainer: a two-metapath model ([[1, 0], [3]],
    embeddings concatenated, model.py:203-220) against the unmodified reference's 5-epoch dropout-free trace and
    test result, and against the epoch-by-epoch Python path (MPNetm + torch.optim.Adam) of this repo."""
    g = load_golden("model_len3")
    data = _bag(fx3)
    metas = [[1, 0], [3]]
    ref, ref_test = g["trace5_m10_m3"], g["trace5_m10_m3_test"]
    torch.manual_seed(30)
    model = mpgnn_b200.MPNetm(2, 64, fx3["num_relations"], 64, 2, len(metas), metas, device="cpu")
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    traces = {}
    for use_graph in (True, False):
        tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, metas, dropout_p=0.0, max_epochs=5)
        tr.load_state_dict(sd0)
        got = tr.state_dict()
        assert list(got) == list(sd0)
        for k, v in sd0.items():
            assert torch.equal(got[k], v), k                      # flat layout round trip incl. the [H, 2H] fc1 transpose
        trace = tr.run(5, use_graph=use_graph)[:5]
        traces[use_graph] = trace
        assert np.allclose(trace[:, 0], ref[:, 0], rtol=1e-4), (trace[:, 0], ref[:, 0])
        assert np.allclose(trace[:, 1], ref[:, 1], rtol=1e-4)
        assert np.allclose(trace[:, 2:], ref[:, 2:], atol=2e-3)
        loss_t, f1_t = tr.evaluate("test")
        assert abs(loss_t - ref_test[0]) < 1e-3 * abs(ref_test[0]) and abs(f1_t - ref_test[1]) < 2e-3
    assert np.array_equal(traces[True], traces[False])
    # the reference-facing call takes the native path for unions too and returns the test macro-F1
    torch.manual_seed(30)
    f1_native = mpgnn_b200.mpgnn_parallel_multiple_x(data, 2, 64, 4, 64, 2, metas, True, epochs=5)
    model.load_state_dict(sd0)
    model.to("cuda")
    model.dropout.p = model.dropout2.p = 0.0
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0005)
    for _ in range(5):
        mpgnn_b200.mpgnn_train(model, opt, data)
    sd5 = tr.state_dict()
    for k, v in model.state_dict().items():
        assert rel_err(sd5[k], v.cpu()) < 1e-4, k
    assert 0.0 <= f1_native <= 1.0


def test_trainer_validate_last_returns_the_every_epoch_result(fx3):
    """MPGNN_TRAINER_VALIDATE_LAST skips the validation passes whose results the reference discards (main.py:1134
    returns the last epoch's): same parameters bit for bit, same returned F1, NaN where nothing was computed."""
    g = load_golden("model_len3")
    data = _bag(fx3)
    out = {}
    for every in (True, False):
        tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, [1, 0], dropout_p=0.6, seed=11, max_epochs=40)
        tr.load_state_dict(_sd(g, "sd0."))
        trace = tr.run(40, validate_every_epoch=every)[:40]
        out[every] = (trace, tr.last_val_f1, tr.state_dict())
    full, last = out[True], out[False]
    assert full[1] == last[1] == full[0][-1, 3]
    assert np.array_equal(full[0][:, 0], last[0][:, 0])                       # same train losses
    assert np.array_equal(full[0][-1], last[0][-1]) and np.isnan(last[0][:-1, 1:]).all()
    for k in full[2]:
        assert torch.equal(full[2][k], last[2][k]), k


def test_trainer_requires_parameters_and_valid_labels(fx3):
    data = _bag(fx3)
    tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, [1, 0], dropout_p=0.0, max_epochs=3)
    with pytest.raises(ValueError):
        tr.run(1)                                       # no load_state_dict yet: refuse to train a recycled slab's bytes
    bad = _bag(fx3)
    bad.train_y = fx3["train_y"].clone()
    bad.train_y[0] = 5
    with pytest.raises(ValueError):
        mpgnn_b200.CandidateTrainer(bad, 2, 64, 2, [1, 0])
    bad2 = _bag(fx3)
    bad2.val_idx = list(fx3["val_idx"])
    bad2.val_idx[3] = fx3["x"].size(0) + 7
    with pytest.raises(ValueError):
        mpgnn_b200.CandidateTrainer(bad2, 2, 64, 2, [1, 0])
    # a relation id the edge list never uses is an empty relation, as in the reference (edge_type == r is empty)
    tr9 = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, [9, 0], dropout_p=0.0, max_epochs=2)
    torch.manual_seed(30)
    tr9.load_state_dict(mpgnn_b200.MPNetm(2, 64, 4, 64, 2, 1, [[9, 0]], device="cpu").state_dict())
    assert np.isfinite(tr9.run(2)[:2]).all()


def test_trainer_union_of_three_metapaths_on_a_large_graph():
    """The final selection trains unions of up to three metapaths (main.py:1465-1475) at the size of the searched graph;
    the scratch of the head's split-K reductions depends on the shape in a non-monotone way, so this runs one at
    400k nodes (where round 2's first sizing overflowed) and checks it against the epoch-by-epoch Python path."""
    gen = torch.Generator().manual_seed(11)
    n, e, r = 400_000, 1_600_000, 6
    ei = torch.randint(0, n, (2, e), generator=gen)
    et = torch.randint(0, r, (e,), generator=gen)
    x = torch.nn.functional.one_hot(torch.randint(0, 2, (n,), generator=gen), 2).float()
    y = torch.randint(0, 2, (n,), generator=gen)
    perm = torch.randperm(n, generator=gen)
    data = mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, num_nodes=n, train_idx=perm[:200_000], train_y=y[perm[:200_000]],
                           val_idx=perm[200_000:300_000], val_y=y[perm[200_000:300_000]], test_idx=perm[300_000:],
                           test_y=y[perm[300_000:]])
    metas = [[3], [1, 3], [5, 2, 0]]
    torch.manual_seed(30)
    model = mpgnn_b200.MPNetm(2, 64, r, 64, 2, len(metas), metas, device="cpu")
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, metas, dropout_p=0.0, max_epochs=3)
    tr.load_state_dict(sd0)
    trace = tr.run(3)[:3]
    assert np.isfinite(trace).all()
    model.to("cuda")
    model.dropout.p = model.dropout2.p = 0.0
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0005)
    for ep in range(3):
        loss, _ = mpgnn_b200.mpgnn_train(model, opt, data)
        assert abs(loss - trace[ep, 0]) < 1e-4 * abs(trace[ep, 0]), (ep, loss, trace[ep, 0])


def test_trainer_compact_hop_equals_dense_hop():
    """The candidate trainer on a sparse multi-relational graph with the hidden layers' aggregated features kept compact
    (what it does by itself from 2^19 nodes on) against the dense form: same losses to fp32 rounding, epoch by epoch,
    CUDA-graph replay included."""
    gen = torch.Generator().manual_seed(21)
    n, r = 60_000, 8
    e = int(0.3 * n * r)
    ei = torch.randint(0, n, (2, e), generator=gen)
    et = torch.randint(0, r, (e,), generator=gen)
    x = torch.nn.functional.one_hot(torch.randint(0, 2, (n,), generator=gen), 2).float()
    y = torch.randint(0, 2, (n,), generator=gen)
    perm = torch.randperm(n, generator=gen)
    data = mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, num_nodes=n, train_idx=perm[:30_000], train_y=y[perm[:30_000]],
                           val_idx=perm[30_000:45_000], val_y=y[perm[30_000:45_000]], test_idx=perm[45_000:],
                           test_y=y[perm[45_000:]])
    metas = [[5, 2, 0], [1, 3]]
    torch.manual_seed(30)
    sd0 = mpgnn_b200.MPNetm(2, 64, r, 64, 2, len(metas), metas, device="cpu").state_dict()
    traces = {}
    for layout in ("dense", "compact"):
        tr = mpgnn_b200.CandidateTrainer(data, 2, 64, 2, metas, dropout_p=0.6, seed=9, max_epochs=6, h_layout=layout)
        tr.load_state_dict(sd0)
        traces[layout] = tr.run(6)[:6]
    assert np.isfinite(traces["compact"]).all()
    assert np.allclose(traces["compact"][:, :2], traces["dense"][:, :2], rtol=1e-4)
    assert np.allclose(traces["compact"][:, 2:], traces["dense"][:, 2:], atol=1e-3)
