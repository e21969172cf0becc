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


# --------------------------------------------------------------------------------------------
# bag iterations of the search (bag mode)
# --------------------------------------------------------------------------------------------
BAG_SEED_BASE = 2000
BAG_EPOCHS = 50              # main.py:890


def bag_seed(metapath_len, relation):
    return BAG_SEED_BASE + 100 * int(metapath_len) + int(relation)


def create_bags(edge_dict, dest_dict):
    """main.py:545-572: per source, the destinations whose every source label is > 0.9 form one
    positive bag (edge order, duplicates kept); every other destination is a singleton negative bag;
    duplicates (by list equality) are dropped keeping the first.  -> (bags, labels)."""
    bag, labels, seen_single = [], [], set()
    for key in edge_dict:
        lst = []
        for value in edge_dict[key]:
            if min(dest_dict[value]) > 0.9:
                lst.append(value)
            elif value not in seen_single and [value] not in bag:
                bag.append([value])
                labels.append(0)
                seen_single.add(value)
        if lst:
            bag.append(lst)
            labels.append(1)
    new_bag, new_labels, seen = [], [], set()
    for b, l in zip(bag, labels):
        t = tuple(b)
        if t not in seen:
            seen.add(t)
            new_bag.append(b)
            new_labels.append(l)
    return new_bag, new_labels


def connected_relations_bags(edge_index, edge_type, bags):
    """main.py:58-67: relations of edges whose row node is in any bag, first-appearance order."""
    ei, et = np.asarray(edge_index), np.asarray(edge_type)
    s = np.array(sorted({v for b in bags for v in b}), dtype=np.int64)
    pos = et[np.isin(ei[0], s)]
    _, first = np.unique(pos, return_index=True)
    return [int(v) for v in pos[np.sort(first)]]


def bag_dictionaries(edge_index, edge_type, relation, bags, bag_labels):
    """create_edge_dictionary(BAGS=True) (main.py:387-407, 426-438): sources = bag nodes in first-
    appearance order; edge_dict over the sources that have an edge of the relation; dest_dict[dst] =
    labels of all bags containing each of its sources, in edge order."""
    mask = []
    seen = set()
    for b in bags:
        for v in b:
            if v not in seen:
                seen.add(v)
                mask.append(v)
    ei = np.asarray(edge_index)
    sel = np.asarray(edge_type) == int(relation)
    rows, cols = ei[0][sel].tolist(), ei[1][sel].tolist()
    present = set(rows)
    edge_dict = {s: [] for s in mask if s in present}
    tmp = {}
    for b, l in zip(bags, bag_labels):
        for v in b:
            tmp.setdefault(v, []).append(float(l))
    dest_dict = {}
    for s, d in zip(rows, cols):
        if s in seen:
            edge_dict[s].append(d)
            dest_dict.setdefault(d, []).extend(tmp[s])
    return mask, edge_dict, dest_dict


def clean_bags_for_relation_type(bags, bag_labels, edge_dict):
    """main.py:579-594: keep, per bag, the nodes that have an edge of the relation; drop empty bags."""
    keep, keep_labels = [], []
    for b, l in zip(bags, bag_labels):
        t = [v for v in b if v in edge_dict]
        if t:
            keep.append(t)
            keep_labels.append(float(l))
    return keep, keep_labels


def _bag_forward(w, lin, x, bags, edge_dict):
    """OutputLayer.forward(BAGS=True) (model.py:45-72).  Returns per-bag prediction, and for the
    gradient the (destination, source) that produced it; plus the per-source values / the per-bag
    argmax map keyed like the reference (str(bag))."""
    pred = torch.zeros(len(bags))
    best_dst, best_src = [-1] * len(bags), [-1] * len(bags)
    src_val, bag_arg = {}, {}
    a_cache = {}
    for i, bag in enumerate(bags):
        cur = -10.0
        for s in bag:
            if s in edge_dict:
                if s not in a_cache:
                    a_cache[s] = float(torch.dot(x[s], lin))            # LinearLayerAttri(feat[s])
                a = np.float32(a_cache[s])
                dsts = edge_dict[s]
                vals = (w[dsts].numpy() * a).astype(np.float32)
                k = int(np.argmax(vals))
                val = np.float32(w[dsts[k]].item()) * a
                src_val[s] = float(val)
                if val > cur:
                    cur = val
                    bag_arg[str(bag)] = dsts[k]
                    pred[i] = float(val)
                    best_dst[i], best_src[i] = dsts[k], s
    return pred, best_dst, best_src, src_val, bag_arg


def score_relation_bags(edge_index, edge_type, relation, x, num_nodes, bags, bag_labels, seed, epochs=BAG_EPOCHS,
                        record=None):
    """score_relation_bags_parallel (main.py:853-917): restarts of `epochs` train() steps in bag mode
    until the loss fails to improve twice; destinations of bags with loss < 1e-4 are frozen (grad
    mask) for the following restarts; the others are re-drawn U(0,1).  Adam(lr=0.1) trains the
    destination weights and the 1 x F `LinearLayerAttri` weight; both are clamped to [0,1].
    -> (current_loss, predictions_for_each_restart, skip flag, final linear weight)."""
    random.seed(seed)
    torch.manual_seed(seed)
    mask, edge_dict, dest_dict = bag_dictionaries(edge_index, edge_type, relation, bags, bag_labels)
    kbags, klabels = clean_bags_for_relation_type(bags, bag_labels, edge_dict)
    y = torch.tensor(klabels, dtype=torch.float32)
    w = initialize_weights(num_nodes, dest_dict)
    grad_mask = torch.ones(num_nodes)
    skip = len(kbags) == 1 or (len(kbags) > 1 and klabels.count(1.0) == 0)
    preds_per_restart, frozen = {}, []
    rest, current_loss = 0, 100.0
    x = torch.as_tensor(x, dtype=torch.float32)
    traj, lin_hist, frozen_hist, w_hist = [], [], [], []
    while rest < 2:
        lin = torch.nn.Linear(x.size(1), 1, bias=False).weight.detach()[0].clone()   # Score.__init__ (model.py:40)
        lin_hist.append(lin.numpy().copy())
        mw, vw = torch.zeros_like(w), torch.zeros_like(w)
        ml, vl = torch.zeros_like(lin), torch.zeros_like(lin)
        for ep in range(1, epochs + 1):
            pred, bd, bs, src_val, bag_arg = _bag_forward(w, lin, x, kbags, edge_dict)
            diff = pred - y
            loss_per_bag = diff * diff
            loss = float(loss_per_bag.mean())
            traj.append(loss)
            gw, gl = torch.zeros_like(w), torch.zeros_like(lin)
            for i in range(len(kbags)):
                if bd[i] >= 0:
                    gi = 2.0 * float(diff[i]) / len(kbags)
                    a = float(torch.dot(x[bs[i]], lin))
                    gw[bd[i]] += gi * a
                    gl += gi * float(w[bd[i]]) * x[bs[i]]
            if frozen:
                gw = gw * grad_mask                                         # main.py:663-664
            bc1, bc2 = 1.0 - 0.9 ** ep, 1.0 - 0.999 ** ep
            for p, g, m_, v_ in ((w, gw, mw, vw), (lin, gl, ml, vl)):
                m_.lerp_(g, 0.1)
                v_.mul_(0.999).addcmul_(g, g, value=0.001)
                p.sub_((SCORER_LR / bc1) * (m_ / ((v_.sqrt() / math.sqrt(bc2)) + 1e-8)))
                p.clamp_(0.0, 1.0)                                          # main.py:667-669
        for k, v in src_val.items():
            preds_per_restart.setdefault(k, []).append(v)
        if loss < current_loss:
            frozen, idx = [], 0
            for _, dst in bag_arg.items():                                  # retrieve_destinations_low_loss (:530-543)
                if float(loss_per_bag[idx]) < 0.0001 and dst not in frozen:
                    frozen.append(dst)
                idx += 1
            current_loss, rest = loss, 0
        else:
            rest += 1
        for node in frozen:
            grad_mask[node] = 0
        frozen_hist.append(list(frozen))
        lin_hist.append(lin.numpy().copy())
        w_hist.append(w.clone())
        nw = torch.zeros(num_nodes)                                         # reinitialize_weights (:499-516)
        fz = set(frozen)
        for key in dest_dict:
            nw[key] = w[key] if key in fz else random.uniform(0.0, 1.0)
        w = nw
    if record is not None:
        record.update(traj=traj, lin_hist=lin_hist, frozen_hist=frozen_hist, w_hist=w_hist, dest_keys=list(dest_dict),
                      bags=kbags, bag_labels=klabels, mask=mask)
    return current_loss, preds_per_restart, skip, lin
