"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the search stage's step-0 relation scorer
(SURVEY §8 a12-a17, a19).  Pinned against tests/golden/search_len3.npz (recorded from the
unmodified reference, generator tests/golden/make_golden_search.py).  Only tests/ and bench.py's
CPU legs may import it.

Randomness: the reference leaves Python's `random` unseeded (main.py:494).  The seam shared by the
reference goldens, this oracle and the product is `random.seed(SCORER_SEED_BASE + relation)` right
before a relation is scored.
"""
import math
import random

import numpy as np
import torch

SCORER_SEED_BASE = 1000
SCORER_EPOCHS = 100          # main.py:755
SCORER_LR = 0.1              # main.py:522


def connected_relations_step0(edge_index, edge_type, labels):
    """node_types_and_connected_relations(BAGS=False, dataset='synthetic') (main.py:69-75): the
    relations of edges whose row node has label 1, in first-appearance (edge) order."""
    ei = np.asarray(edge_index)
    et = np.asarray(edge_type)
    lab = np.asarray(labels).reshape(-1)
    pos = et[lab[ei[0]] == 1]
    _, first = np.unique(pos, return_index=True)
    return [int(v) for v in pos[np.sort(first)]]


def relation_dictionaries(edge_index, edge_type, relation, labels):
    """score_relation_parallel's setup (main.py:733-737 -> create_edge_dictionary :387-425) for the
    first iteration: sources = sorted unique row nodes of the relation; edge_dict {src: [dst...]} in
    source order with destinations in edge order (duplicates kept); dest_dict {dst: [label of each of
    its sources...]} keyed in first-appearance edge order."""
    ei = np.asarray(edge_index)
    sel = np.asarray(edge_type) == int(relation)
    rows, cols = ei[0][sel], ei[1][sel]
    lab = np.asarray(labels).reshape(-1)
    sources = np.unique(rows)
    edge_dict = {int(s): [] for s in sources}
    dest_dict = {}
    for s, d in zip(rows.tolist(), cols.tolist()):
        edge_dict[s].append(d)
        dest_dict.setdefault(d, []).append(int(lab[s]))
    return edge_dict, dest_dict


def initialize_weights(num_nodes, dest_dict, rng=random):
    """main.py:479-497: w[dst] = |min(labels of its sources) + U(-0.2, 0.2)| in dict order (entries
    of nodes that are not destinations are uninitialised memory in the reference and never read;
    0 here)."""
    w = torch.zeros(num_nodes)
    for key, values in dest_dict.items():
        w[key] = abs(min(values) + rng.uniform(-0.2, 0.2))
    return w


def score_relation(edge_index, edge_type, relation, labels, num_nodes, epochs=SCORER_EPOCHS, seed_base=SCORER_SEED_BASE):
    """score_relation_parallel (main.py:727-760) + train (:641-673) + Score/OutputLayer forward
    (model.py:75-88), non-bag mode: pred[src] = max_dst w[dst] (first maximum), MSE(mean) against the
    source labels, Adam(lr=0.1) on w, clamp to [0,1].  Returns (loss trajectory, final w, argmax dst
    per source of the last forward, dest keys)."""
    random.seed(seed_base + int(relation))
    edge_dict, dest_dict = relation_dictionaries(edge_index, edge_type, relation, labels)
    w = initialize_weights(num_nodes, dest_dict).double().float()
    sources = list(edge_dict.keys())
    lab = torch.as_tensor(np.asarray(labels).reshape(-1), dtype=torch.float32)
    y = lab[torch.tensor(sources)]
    m = torch.zeros_like(w)
    v = torch.zeros_like(w)
    traj, arg = [], None
    for ep in range(1, epochs + 1):
        pred = torch.empty(len(sources))
        arg = []
        for i, s in enumerate(sources):
            dsts = edge_dict[s]
            k = int(torch.argmax(w[dsts]))
            arg.append(dsts[k])
            pred[i] = w[dsts[k]]
        diff = pred - y
        traj.append(float((diff * diff).mean()))
        g = torch.zeros_like(w)
        g.index_add_(0, torch.tensor(arg), 2.0 * diff / len(sources))
        # torch.optim.Adam(lr=0.1), defaults otherwise
        m.lerp_(g, 1.0 - 0.9)
        v.mul_(0.999).addcmul_(g, g, value=1.0 - 0.999)
        bc1, bc2 = 1.0 - 0.9 ** ep, 1.0 - 0.999 ** ep
        w = w - (SCORER_LR / bc1) * (m / ((v.sqrt() / math.sqrt(bc2)) + 1e-8))
        w = w.clamp(0.0, 1.0)                                   # main.py:667
    return traj, w, arg, list(dest_dict.keys())


def gap_select_step0(relations, losses):
    """main.py:1346-1355: sort the losses, take the largest gap, keep relations with loss <= the
    value just below it; keep everything when there are fewer than two gaps."""
    accs = sorted(losses)
    diffs = np.diff(accs)
    if len(diffs) >= 2:
        idx = int(np.argmax(diffs))
        return [r for r, l in zip(relations, losses) if l <= accs[idx]]
    return list(relations)


def candidate_blocks(n_items, size, rank):
    """main.py:1444-1450: contiguous block partition, the first n_items % size ranks get one more."""
    sub, rem = n_items // size, n_items % size
    start = rank * sub + min(rank, rem)
    return start, start + sub + (1 if rank < rem else 0)


def relation_split(relations, size, rank):
    """main.py:1319: np.array_split(actual_relations, size)[rank]."""
    return [int(v) for v in np.array_split(np.asarray(relations), size)[rank]]
