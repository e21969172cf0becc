"""Input boundary of the path: the reference's TSV loaders and split, restated.

File formats (written by the reference's generator / preprocessing scripts):
  node.dat   `id \\t f0 \\t f1 ...`      link.dat   `src \\t relation \\t dst`      label.dat  `id \\t label`
`src` (column 0 / edge_index[0]) is the aggregation target, `dst` the message source.
"""
import numpy as np
import pandas as pd
import torch
from sklearn.model_selection import train_test_split


def load_files(node_file_path, link_file_path, label_file_path):
    """main.py:178-195 -> (labels, features, links, [labels], tot_relation_types)."""
    features = pd.read_csv(node_file_path, sep="\t", header=None)
    features = features.dropna(axis=1, how="all")
    features.rename(columns={0: "node", 1: "features"}, inplace=True)
    labels_df = pd.read_csv(label_file_path, sep="\t", header=None)
    labels_df.rename(columns={0: "node", 1: "label"}, inplace=True)
    labels = torch.tensor(labels_df["label"].values)
    links = pd.read_csv(link_file_path, sep="\t", header=None)
    links.rename(columns={0: "node_1", 1: "relation_type", 2: "node_2"}, inplace=True)
    tot_relation_types = len(set(links["relation_type"].to_list()))
    return labels, features, links, [labels], tot_relation_types


def load_files_fb15k237(node_file_path, link_file_path, label_file_path, relations_legend_path=None):
    """main.py:138-176 -> (labels, features, links, labelled source nodes, tot_relation_types,
    one-vs-rest binary label sets)."""
    labels, features, links, _, tot = load_files(node_file_path, link_file_path, label_file_path)
    labels_df = pd.read_csv(label_file_path, sep="\t", header=None)
    sources = labels_df[0].values.tolist()
    uniq = torch.unique(labels).tolist()
    if len(uniq) > 2:
        binary = [(labels == u).to(labels.dtype) for u in uniq]
    else:
        binary = [labels]
    return labels, features, links, sources, tot, binary


def get_node_features(colors):
    """main.py:347-355: one-hot of the non-numeric columns (pd.get_dummies), `node` dropped, fp32."""
    node_features = pd.get_dummies(colors)
    node_features = node_features.drop(["node"], axis=1)
    return torch.from_numpy(node_features.to_numpy().astype(np.float32))


def sn(test_index, val_index, train_index, feature_matrix):
    """main.py:357-364: zero the features of every labelled node (fb15k-237 only)."""
    idx = torch.as_tensor(list(test_index) + list(val_index) + list(train_index), dtype=torch.long)
    feature_matrix[idx] = 0
    return feature_matrix


def get_edge_index_and_type_no_reverse(links):
    """main.py:366-372 -> (edge_index int64 [2,E] = [node_1; node_2], edge_type int64 [E])."""
    edge_index = torch.tensor(np.stack([links["node_1"].values, links["node_2"].values]).astype(np.int64))
    edge_type = torch.tensor(links["relation_type"].values.astype(np.int64))
    return edge_index, edge_type


def splitting_node_and_labels(lab, feat, src, dataset):
    """main.py:277-345: two stratified sklearn splits (random_state=415; 10 % test, then 20 % of the
    rest validation); classes with a single member are held out and appended to the training set."""
    node_idx = list(feat["node"].values) if dataset == "synthetic" else list(src)
    lab_list = lab.tolist()
    counts = {}
    for i, v in enumerate(lab_list):
        counts.setdefault(v, []).append(i)
    unique_indices = [ix[0] for ix in counts.values() if len(ix) == 1]
    removed_nodes, removed_lab = [], []
    if unique_indices:
        for i in sorted(unique_indices, reverse=True):
            removed_nodes.append(node_idx.pop(i))
            removed_lab.append(lab_list.pop(i))
        lab_used = lab_list
    else:
        lab_used = lab
    train_idx, test_idx, train_y, test_y = train_test_split(node_idx, lab_used, random_state=415, stratify=lab_used,
                                                            test_size=0.1)
    train_idx, val_idx, train_y, val_y = train_test_split(train_idx, train_y, random_state=415, stratify=train_y,
                                                          test_size=0.2)
    if unique_indices:
        train_idx.extend(removed_nodes)
        train_y.extend(removed_lab)
        return (torch.tensor(node_idx), train_idx, torch.tensor(train_y), test_idx, torch.tensor(test_y), val_idx,
                torch.tensor(val_y))
    return torch.tensor(node_idx), train_idx, train_y, test_idx, test_y, val_idx, val_y
