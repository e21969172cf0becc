"""Loaders and split round-tripped through the reference's TSV formats and checked
against the arrays the reference's own loaders produced (tests/golden/fixture_len3.npz, which also
holds the reference's labels.pkl split)."""
import numpy as np
import torch

from conftest import load_golden

from mpgnn_b200 import data as D


def _write_tsv(tmp_path, g):
    x, ei, et, lab = g["x"], g["edge_index"], g["edge_type"], g["labels"]
    node = tmp_path / "node.dat"
    link = tmp_path / "link.dat"
    label = tmp_path / "label.dat"
    with open(node, "w") as f:
        for i, row in enumerate(x):
            f.write("%d\t%s\n" % (i, "\t".join(str(int(v)) for v in row)))
    with open(link, "w") as f:
        for s, r, d in zip(ei[0], et, ei[1]):
            f.write("%d\t%d\t%d\n" % (s, r, d))
    with open(label, "w") as f:
        for i, v in enumerate(lab):
            f.write("%d\t%d\n" % (i, v))
    return str(node), str(link), str(label)


def test_loaders_and_split_match_reference(tmp_path):
    g = load_golden("fixture_len3")
    node, link, label = _write_tsv(tmp_path, g)
    labels, features, links, binary, n_rel = D.load_files(node, link, label)
    assert n_rel == int(g["num_relations"]) and len(binary) == 1
    assert np.array_equal(labels.numpy(), g["labels"])
    x = D.get_node_features(features)
    assert x.dtype == torch.float32 and np.array_equal(x.numpy(), g["x"])
    ei, et = D.get_edge_index_and_type_no_reverse(links)
    assert ei.dtype == torch.int64 and np.array_equal(ei.numpy(), g["edge_index"])
    assert np.array_equal(et.numpy(), g["edge_type"])
    node_idx, tr_i, tr_y, te_i, te_y, va_i, va_y = D.splitting_node_and_labels(labels, features, [], "synthetic")
    # the reference's labels.pkl fixture pins the split, order included
    assert [int(v) for v in tr_i] == g["labels_pkl_train"][:, 0].tolist()
    assert [int(v) for v in tr_y] == g["labels_pkl_train"][:, 1].tolist()
    assert [int(v) for v in va_i] == g["labels_pkl_val"][:, 0].tolist()
    assert [int(v) for v in te_i] == g["labels_pkl_test"][:, 0].tolist()
    assert [int(v) for v in te_y] == g["labels_pkl_test"][:, 1].tolist()


def test_sn_zeroes_labelled_rows():
    x = torch.ones(6, 3)
    out = D.sn([0], [2], [5], x)
    assert out.sum(1).tolist() == [0, 3, 0, 3, 3, 0]


def test_rewritten_fixture_files_round_trip_through_the_loaders(tmp_path):
    """tests/test_gpu_cli.py feeds the CLI with TSV files rewritten from the committed golden (the reference's folder
    does not travel to the GPU box): loading them must give back the golden's tensors exactly."""
    from conftest import write_fixture_files, fixture_as_torch
    from mpgnn_b200 import data as D
    fx3 = fixture_as_torch("fixture_len3")
    folder = write_fixture_files(str(tmp_path / "fx"))
    labels, features, links, binary, tot = D.load_files(folder + "/node.dat", folder + "/link.dat", folder + "/label.dat")
    assert torch.equal(D.get_node_features(features), fx3["x"])
    ei, et = D.get_edge_index_and_type_no_reverse(links)
    assert torch.equal(ei, fx3["edge_index"]) and torch.equal(et, fx3["edge_type"]) and tot == fx3["num_relations"]
    assert torch.equal(labels, fx3["labels"])
    _, train_idx, train_y, test_idx, test_y, val_idx, val_y = D.splitting_node_and_labels(labels, features, [], "synthetic")
    assert list(train_idx) == fx3["train_idx"] and list(val_idx) == fx3["val_idx"] and list(test_idx) == fx3["test_idx"]
