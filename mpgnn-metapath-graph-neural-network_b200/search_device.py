"""The bag pipeline between two search steps as tensor set-operations on the device.

`search.py` restates the reference's dictionary/list functions one by one under their reference names
(create_edge_dictionary, create_bags, clean_bags_for_relation_type, relabel_nodes_inside_bags, clean_dictionaries,
node_types_and_connected_relations; main.py:56-84, 387-477, 545-634).  They are host Python: at the configs[4] size
(1M nodes x 100 relations) a relation costs ~0.1 s of dictionary building next to 3.5 ms of scoring on the device.
This module keeps the same state as ARRAYS on the device and derives every step with sorts / uniques / scatters:

  reference object                          array form
  edge_dictionary {src: [dst...]}           `src_order` (keys in dict order) + the relation's CSR slice of each key (edge order)
  destination_dictionary {dst: [labels]}    `dst_keys` (first-appearance order) + per-node counts of 0- and 1-labels
                                            (every consumer takes min(), `0 in list`, or removes one 0: counts suffice)
  data.bags / data.bag_labels               ragged (`bag_ptr`, `bag_flat`) + labels, in the reference's append order
  predictions_for_each_restart {src: [..]}  visited order + a [restarts, N] matrix

Orders are the reference's (first appearance in edge order, dict insertion order, `later duplicates overwrite`): they
decide the order of the `random.uniform` draws and therefore every number downstream.  The draws themselves come from
Python's own Mersenne Twister state, advanced in bulk through numpy (same generator, same doubles), so the stream a
relation consumes under `random.seed(seed)` is bit for bit the one the dictionary path consumes.

Everything here is torch on `device` (CUDA in production; the CPU tests run the same code against the dictionary
functions).  Only the two K5 calls (`search.run_scorer`, `run_bag_restart_arrays`) need the CUDA library.
"""
import random

import numpy as np
import torch

from . import _lib

BAG_EPOCHS, SCORER_LR = 50, 0.1          # main.py:890, 522


# ---------------------------------------------------------------------------------------------------------------------
# Python's `random` stream in bulk
# ---------------------------------------------------------------------------------------------------------------------
def mt_draws(count):
    """`count` consecutive `random.random()` values as a float64 array, advancing Python's global generator exactly as
    `count` calls would (random.random and numpy's RandomState.random_sample are the same MT19937 genrand_res53)."""
    if count <= 0:
        return np.zeros(0, dtype=np.float64)
    st = random.getstate()
    rs = np.random.RandomState()
    rs.set_state(("MT19937", np.array(st[1][:-1], dtype=np.uint32), st[1][-1]))
    out = rs.random_sample(int(count))
    s = rs.get_state()
    random.setstate((st[0], tuple(int(v) for v in s[1]) + (int(s[2]),), st[2]))
    return out


def uniform_draws(lo, hi, count):
    """[random.uniform(lo, hi) for _ in range(count)]: lo + (hi - lo) * random() in double, as CPython computes it."""
    return lo + (hi - lo) * mt_draws(count)


# ---------------------------------------------------------------------------------------------------------------------
# small tensor helpers
# ---------------------------------------------------------------------------------------------------------------------
def first_unique(values):
    """Distinct entries of a 1-D int tensor in order of first appearance (a dict's key order)."""
    if values.numel() == 0:
        return values
    u, inv = torch.unique(values, return_inverse=True)
    first = torch.full((u.numel(),), values.numel(), dtype=torch.int64, device=values.device)
    first.scatter_reduce_(0, inv, torch.arange(values.numel(), device=values.device), reduce="amin", include_self=True)
    return u[torch.argsort(first)]


def ragged_arange(starts, lens):
    """Concatenation of arange(starts[i], starts[i] + lens[i]) and the segment id of every element."""
    total = int(lens.sum())
    seg = torch.repeat_interleave(torch.arange(lens.numel(), device=lens.device), lens, output_size=total)
    excl = torch.cumsum(lens, 0) - lens
    return starts[seg] + (torch.arange(total, device=lens.device) - excl[seg]), seg


def _mix(a, b, salt):
    """64-bit mixing of (a, b) element-wise (splitmix-style, wrap-around int64 arithmetic)."""
    z = a * -7046029254386353131 + b * -4658895280553007687 + salt
    z = (z ^ (z >> 30)) * -4658895280553007687
    z = (z ^ (z >> 27)) * -7723592293110705685
    return z ^ (z >> 31)


def ragged_hash(flat, seg, n_seg):
    """Two independent 64-bit content hashes per ragged list (order- and multiplicity-sensitive)."""
    start = torch.zeros(n_seg, dtype=torch.int64, device=flat.device)
    lens = torch.zeros(n_seg, dtype=torch.int64, device=flat.device).scatter_add_(0, seg, torch.ones_like(seg))
    start = torch.cumsum(lens, 0) - lens
    pos = torch.arange(flat.numel(), device=flat.device) - start[seg]
    h1 = torch.zeros(n_seg, dtype=torch.int64, device=flat.device).scatter_add_(0, seg, _mix(flat, pos, 0x243F6A8885A308D3))
    h2 = torch.zeros(n_seg, dtype=torch.int64, device=flat.device).scatter_add_(0, seg, _mix(pos, flat, 0x13198A2E03707344))
    return torch.stack([lens, h1, h2], dim=1)


def first_occurrence(keys2d):
    """For rows of an int matrix: (index of the first row equal to row i, for every i; mask of the first occurrences)."""
    if keys2d.size(0) == 0:
        return torch.zeros(0, dtype=torch.int64, device=keys2d.device), torch.zeros(0, dtype=torch.bool, device=keys2d.device)
    _, inv = torch.unique(keys2d, dim=0, return_inverse=True)
    n = keys2d.size(0)
    first = torch.full((int(inv.max()) + 1,), n, dtype=torch.int64, device=keys2d.device)
    first.scatter_reduce_(0, inv, torch.arange(n, device=keys2d.device), reduce="amin", include_self=True)
    rep = first[inv]
    return rep, rep == torch.arange(n, device=keys2d.device)


# ---------------------------------------------------------------------------------------------------------------------
# graph side
# ---------------------------------------------------------------------------------------------------------------------
class SearchGraph:
    """The edge list on the device, bucketed per relation in ORIGINAL edge order (what `edge_index[:, edge_type == r]`
    enumerates, mp_rgcn_layer.py:29-37), plus the per-relation CSR by source (a source's destinations in edge order =
    its edge_dictionary entry)."""

    def __init__(self, edge_index, edge_type, num_nodes, device):
        self.device = torch.device(device)
        self.n = int(num_nodes)
        ei = torch.as_tensor(edge_index).to(self.device, torch.int64)
        self.et = torch.as_tensor(edge_type).to(self.device, torch.int64)
        self.rows_all, self.cols_all = ei[0].contiguous(), ei[1].contiguous()
        self.r = int(self.et.max()) + 1 if self.et.numel() else 1
        order = torch.argsort(self.et, stable=True)
        self.rows, self.cols = self.rows_all[order], self.cols_all[order]
        self.rel_ptr = [0] + torch.cumsum(torch.bincount(self.et, minlength=self.r), 0).tolist()
        self._csr = {}

    def rel_edges(self, rel):
        if rel >= self.r:
            z = torch.zeros(0, dtype=torch.int64, device=self.device)
            return z, z
        a, b = self.rel_ptr[rel], self.rel_ptr[rel + 1]
        return self.rows[a:b], self.cols[a:b]

    def csr(self, rel):
        """(ptr [N+1], dst) of relation `rel` bucketed by source, destinations of a source in edge order."""
        hit = self._csr.get(rel)
        if hit is None:
            rows, cols = self.rel_edges(rel)
            order = torch.argsort(rows, stable=True)
            ptr = torch.zeros(self.n + 1, dtype=torch.int64, device=self.device)
            ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=self.n), 0)
            hit = (ptr, cols[order])
            if len(self._csr) >= 4:
                self._csr.pop(next(iter(self._csr)))
            self._csr[rel] = hit
        return hit

    def connected_relations(self, keep_rows_mask):
        """node_types_and_connected_relations (main.py:56-84): relations of the edges whose source passes the node mask,
        in first-appearance edge order."""
        return self.connected_relations_from_edge_mask(keep_rows_mask[self.rows_all])

    def connected_relations_from_edge_mask(self, keep_edges):
        return [int(v) for v in first_unique(self.et[keep_edges]).tolist()]


class BagState:
    """What current_metapaths_dict[str(metapath)] holds in the reference (main.py:1365-1369, 1435): the dictionaries of
    the metapath's first relation restricted to its current source set, and the Data copy with labels / bags."""

    def __init__(self, rel, src_order, count0, count1, labels):
        self.rel = int(rel)
        self.src_order = src_order          # edge_dictionary keys, dict order
        self.count0, self.count1 = count0, count1      # per destination node: how many 0- / 1-labels its list holds
        self.labels = labels                # data.labels as a float [N] tensor
        self.bag_ptr = self.bag_flat = self.bag_labels = None

    def copy(self):
        new = BagState(self.rel, self.src_order, self.count0, self.count1, self.labels)
        new.bag_ptr, new.bag_flat, new.bag_labels = self.bag_ptr, self.bag_flat, self.bag_labels
        return new

    def bags_as_lists(self):
        ptr, flat = self.bag_ptr.tolist(), self.bag_flat.tolist()
        return [flat[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]


def _source_labels(sg, rows, labels, source_list, dataset):
    """Label the reference attaches to an edge's source (main.py:421-424): the node's own label for 'synthetic', the
    label at the source's POSITION in the source list otherwise."""
    if dataset == "synthetic":
        return labels[rows]
    pos = torch.full((sg.n,), -1, dtype=torch.int64, device=sg.device)
    src = torch.as_tensor(source_list, dtype=torch.int64, device=sg.device)
    # list.index(): the first position of a repeated node
    pos.scatter_reduce_(0, src, torch.arange(src.numel(), device=sg.device), reduce="amin", include_self=False)
    return labels[pos[rows]]


def step0_inputs(sg, rel, labels, source_list, dataset):
    """score_relation_parallel's setup (main.py:727-743) without dictionaries -> (initial weights [N] float32 on the
    device, source mask uint8 [N] or None, node labels float32 [N] for the scorer)."""
    rows, cols = sg.rel_edges(rel)
    first = not source_list
    if first:
        in_mask = torch.ones(rows.numel(), dtype=torch.bool, device=sg.device)
    else:
        member = torch.zeros(sg.n, dtype=torch.bool, device=sg.device)
        member[torch.as_tensor(source_list, dtype=torch.int64, device=sg.device)] = True
        in_mask = member[rows]
    rows_m, cols_m = rows[in_mask], cols[in_mask]
    src_lab = _source_labels(sg, rows_m, labels, source_list, dataset).to(torch.float64)
    keys = first_unique(cols_m)
    weights = torch.zeros(sg.n, dtype=torch.float32, device=sg.device)
    if keys.numel():
        mins = torch.full((sg.n,), float("inf"), dtype=torch.float64, device=sg.device)
        mins.scatter_reduce_(0, cols_m, src_lab, reduce="amin", include_self=True)
        draws = torch.from_numpy(uniform_draws(-0.2, 0.2, keys.numel())).to(sg.device)
        weights[keys] = (mins[keys] + draws).abs().to(torch.float32)          # abs(min + U) in double, rounded once
    if dataset == "synthetic":
        node_labels = labels.to(torch.float32)
        mask = None if first else member.to(torch.uint8)
    else:
        node_labels = torch.zeros(sg.n, dtype=torch.float32, device=sg.device)
        src = torch.as_tensor(source_list, dtype=torch.int64, device=sg.device)
        node_labels[src] = labels[:src.numel()].to(torch.float32)
        mask = member.to(torch.uint8)
    return weights, mask, node_labels


def step0_state(sg, rel, labels, source_list, dataset):
    """The dictionaries score_relation_parallel returns next to the loss (main.py:737), as a BagState."""
    rows, cols = sg.rel_edges(rel)
    if not source_list:
        src_order = torch.unique(rows)                                        # torch.unique(...).tolist(): ascending
        in_mask = torch.ones(rows.numel(), dtype=torch.bool, device=sg.device)
    else:
        member = torch.zeros(sg.n, dtype=torch.bool, device=sg.device)
        src = torch.as_tensor(source_list, dtype=torch.int64, device=sg.device)
        member[src] = True
        in_mask = member[rows]
        has = torch.zeros(sg.n, dtype=torch.bool, device=sg.device)
        has[rows] = True
        src_order = first_unique(src[has[src]])                               # keys in the list's order (a repeat is one key)
    return _state_from_edges(sg, rel, src_order, rows[in_mask], cols[in_mask],
                             _source_labels(sg, rows[in_mask], labels, source_list, dataset), labels)


def _state_from_edges(sg, rel, src_order, rows_m, cols_m, src_lab, labels):
    ones = torch.ones(rows_m.numel(), dtype=torch.int32, device=sg.device)
    is1 = (src_lab != 0)
    count0 = torch.zeros(sg.n, dtype=torch.int32, device=sg.device).scatter_add_(0, cols_m[~is1], ones[~is1])
    count1 = torch.zeros(sg.n, dtype=torch.int32, device=sg.device).scatter_add_(0, cols_m[is1], ones[is1])
    return BagState(rel, src_order, count0, count1, labels)


# ---------------------------------------------------------------------------------------------------------------------
# create_bags (main.py:545-572)
# ---------------------------------------------------------------------------------------------------------------------
def create_bags(sg, state):
    """Per source (dict order), in edge order: a destination whose every source label is > 0.9 joins the source's
    positive bag, any other destination becomes a singleton negative bag the first time it is met; the positive bag is
    appended after the source's singletons; bags equal in content to an earlier bag are dropped."""
    ptr, dst = sg.csr(state.rel)
    dev = sg.device
    src = state.src_order
    lens = ptr[src + 1] - ptr[src]
    pos_idx, seg = ragged_arange(ptr[src], lens)
    v = dst[pos_idx]
    positive = (state.count0[v] == 0) & (state.count1[v] > 0)                  # min(labels) > 0.9 on 0/1 labels
    t = torch.arange(v.numel(), device=dev)
    # singletons: first time a non-positive destination is met in the traversal
    neg_v, neg_t = v[~positive], t[~positive]
    if neg_v.numel():
        u, inv = torch.unique(neg_v, return_inverse=True)
        first_t = torch.full((u.numel(),), v.numel(), dtype=torch.int64, device=dev)
        first_t.scatter_reduce_(0, inv, neg_t, reduce="amin", include_self=True)
        single_v, single_key = u, 2 * first_t                                  # key: traversal position (even)
    else:
        single_v = single_key = torch.zeros(0, dtype=torch.int64, device=dev)
    # positive bags: one per source that has any, placed after the source's last edge
    pos_v, pos_seg = v[positive], seg[positive]
    n_src = src.numel()
    pcount = torch.zeros(n_src, dtype=torch.int64, device=dev).scatter_add_(0, pos_seg, torch.ones_like(pos_seg))
    has_pos = pcount > 0
    end_t = torch.cumsum(lens, 0)                                              # one past the source's last traversal slot
    pbag_id = torch.cumsum(has_pos.to(torch.int64), 0) - 1                     # index among the positive bags
    n_pos = int(has_pos.sum())
    if n_pos:
        keys3 = ragged_hash(pos_v, pbag_id[pos_seg], n_pos)
        _, is_first = first_occurrence(keys3)                                  # drop positive bags equal to an earlier one
        pos_key = (2 * end_t[has_pos] - 1)[is_first]                           # odd: after the singletons of the same source
        keep_elem = is_first[pbag_id[pos_seg]]
        pos_v, pos_owner = pos_v[keep_elem], pbag_id[pos_seg][keep_elem]
        plen = pcount[has_pos][is_first]
    else:
        pos_key = plen = torch.zeros(0, dtype=torch.int64, device=dev)
        pos_owner = torch.zeros(0, dtype=torch.int64, device=dev)
    # merge by key
    all_key = torch.cat([single_key, pos_key])
    all_len = torch.cat([torch.ones_like(single_key), plen])
    all_lab = torch.cat([torch.zeros(single_key.numel(), device=dev), torch.ones(pos_key.numel(), device=dev)])
    order = torch.argsort(all_key)
    bag_len = all_len[order]
    bag_ptr = torch.zeros(order.numel() + 1, dtype=torch.int64, device=dev)
    bag_ptr[1:] = torch.cumsum(bag_len, 0)
    # flat members: singletons carry their value; positive bags their (already ordered) elements
    flat = torch.empty(int(bag_ptr[-1]), dtype=torch.int64, device=dev)
    slot_of = torch.empty_like(order)
    slot_of[order] = torch.arange(order.numel(), device=dev)                   # final position of every candidate bag
    n_single = single_key.numel()
    flat[bag_ptr[slot_of[:n_single]]] = single_v
    if pos_key.numel():
        # dense re-index of the kept positive bags, then the position of each element inside its bag
        kept_ids = torch.unique(pos_owner)                                     # ascending == order of pos_key
        dense = torch.searchsorted(kept_ids, pos_owner)
        start = torch.cumsum(plen, 0) - plen
        inner = torch.arange(pos_v.numel(), device=dev) - start[dense]
        flat[bag_ptr[slot_of[n_single + dense]] + inner] = pos_v
    state.bag_ptr, state.bag_flat, state.bag_labels = bag_ptr, flat, all_lab[order].to(torch.float32)
    return state


def bag_member_mask(sg, state):
    m = torch.zeros(sg.n, dtype=torch.bool, device=sg.device)
    m[state.bag_flat] = True
    return m


# ---------------------------------------------------------------------------------------------------------------------
# bag-mode scoring (main.py:853-917) on arrays
# ---------------------------------------------------------------------------------------------------------------------
class BagProblem:
    """Everything score_relation_bags_parallel derives from (data with bags, relation) before it trains."""

    def __init__(self, sg, state, rel):
        dev = sg.device
        ptr, _ = sg.csr(rel)
        has_edge = (ptr[1:] - ptr[:-1]) > 0
        flat = state.bag_flat
        n_bags = state.bag_ptr.numel() - 1
        bag_of = torch.repeat_interleave(torch.arange(n_bags, device=dev), state.bag_ptr[1:] - state.bag_ptr[:-1],
                                         output_size=flat.numel())
        keep = has_edge[flat]                                                  # clean_bags_for_relation_type (:579-594)
        cnt = torch.zeros(n_bags, dtype=torch.int64, device=dev).scatter_add_(0, bag_of[keep], torch.ones_like(bag_of[keep]))
        alive = cnt > 0
        self.bag_flat = flat[keep]
        new_id = torch.cumsum(alive.to(torch.int64), 0) - 1
        self.bag_of = new_id[bag_of[keep]]
        self.bag_ptr = torch.zeros(int(alive.sum()) + 1, dtype=torch.int64, device=dev)
        self.bag_ptr[1:] = torch.cumsum(cnt[alive], 0)
        self.bag_labels = state.bag_labels[alive]
        self.n_bags = int(alive.sum())
        # min label over ALL bags that contain a node (the labels a source hands to its destinations, :431-437)
        srcmin = torch.full((sg.n,), float("inf"), dtype=torch.float64, device=dev)
        srcmin.scatter_reduce_(0, flat, state.bag_labels.to(torch.float64)[bag_of], reduce="amin", include_self=True)
        member = torch.isfinite(srcmin)
        rows, cols = sg.rel_edges(rel)
        sel = member[rows]
        cols_s = cols[sel]
        self.dst_keys = first_unique(cols_s)                                   # destination_bag_dictionary key order
        dmin = torch.full((sg.n,), float("inf"), dtype=torch.float64, device=dev)
        dmin.scatter_reduce_(0, cols_s, srcmin[rows[sel]], reduce="amin", include_self=True)
        self.dst_min = dmin
        self.visited = first_unique(self.bag_flat)                             # key order of max_destination_node_for_source
        self.skip = self.n_bags == 1 or (self.n_bags > 1 and not bool((self.bag_labels == 1).any()))
        # `arg[str(bag)] = dst`: a later bag with the same content overwrites the earlier entry but keeps its slot
        if self.n_bags:
            rep, is_first = first_occurrence(ragged_hash(self.bag_flat, self.bag_of, self.n_bags))
            self.slot_first = torch.nonzero(is_first).reshape(-1)              # dict slots in insertion order = first occurrences
            last = torch.zeros(self.n_bags, dtype=torch.int64, device=dev)
            last.scatter_reduce_(0, rep, torch.arange(self.n_bags, device=dev), reduce="amax", include_self=True)
            self.slot_last = last[self.slot_first]                             # the bag whose destination the slot ends up holding


def run_bag_restart_arrays(graph, rel, prob, x_dev, weights, lin, grad_mask, use_mask, epochs=BAG_EPOCHS, lr=SCORER_LR):
    """One restart on the K5 bag kernel with device-resident inputs and outputs -> (trajectory [epochs] (device),
    trained weights [N], linear weight [F], best destination per bag, (prediction - label) per bag, value per source [N])."""
    lib = _lib.load()
    dev = graph.device
    x_dev = x_dev.contiguous()          # the kernel reads row-major [N, F] (a host array may arrive column-major)
    n, feat = graph.num_nodes, x_dev.size(1)
    nb = prob.n_bags
    src = prob.bag_flat.to(torch.int32).contiguous()
    ptr = prob.bag_ptr.to(torch.int32).contiguous()
    lab = prob.bag_labels.to(torch.float32).contiguous()
    w = weights.to(torch.float32).contiguous().clone()
    ln = lin.to(device=dev, dtype=torch.float32).contiguous().clone()
    gm = grad_mask.to(torch.uint8).contiguous()
    traj = torch.empty(epochs, device=dev)
    best_dst = torch.empty(nb, dtype=torch.int32, device=dev)
    best_src = torch.empty(nb, dtype=torch.int32, device=dev)
    diff = torch.empty(nb, device=dev)
    src_val = torch.full((n,), float("nan"), device=dev)
    ws = torch.empty(lib.mpgnn_score_bags_workspace_bytes(n, nb, feat), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mpgnn_score_bags(graph.handle, int(rel), _lib.ptr(ptr), _lib.ptr(src), nb, _lib.ptr(lab),
                                        _lib.ptr(x_dev), feat, _lib.ptr(w), _lib.ptr(ln), _lib.ptr(gm), int(use_mask),
                                        int(epochs), float(lr), _lib.ptr(traj), _lib.ptr(best_dst), _lib.ptr(best_src),
                                        _lib.ptr(diff), _lib.ptr(src_val), _lib.ptr(ws), ws.numel(), _lib.current_stream()))
    return traj, w, ln, best_dst.to(torch.int64), diff, src_val


def bag_restart_loop(sg, graph, state, rel, x_dev, features_dim, seed, max_restarts=None, restart_fn=None, record=None):
    """score_relation_bags_parallel (restarts until two non-improvements, freezing; main.py:884-911) and, with
    max_restarts=1, retrain_bags (main.py:814-850) -> (best loss, final linear weight [F] (host), per-restart values of
    every visited source as a [restarts, N] device matrix, visited order, skip flag)."""
    restart_fn = restart_fn or run_bag_restart_arrays
    dev = sg.device
    random.seed(seed)
    torch.manual_seed(seed)
    prob = BagProblem(sg, state, rel)
    keys = prob.dst_keys
    weights = torch.zeros(sg.n, dtype=torch.float32, device=dev)
    if keys.numel():                                                          # initialize_weights (:479-497), dict order
        draws = torch.from_numpy(uniform_draws(-0.2, 0.2, keys.numel())).to(dev)
        weights[keys] = (prob.dst_min[keys] + draws).abs().to(torch.float32)
    grad_mask = torch.ones(sg.n, dtype=torch.uint8, device=dev)
    frozen = torch.zeros(sg.n, dtype=torch.bool, device=dev)
    any_frozen = False
    values, lin = [], None
    rest, current_loss, restarts = 0, 100.0, 0
    while rest < 2 and prob.n_bags > 0:
        lin0 = torch.nn.Linear(int(features_dim), 1, bias=False).weight.detach()[0].clone()      # Score.__init__
        traj, trained_w, lin_d, best_dst, diff, src_val = restart_fn(graph, rel, prob, x_dev, weights, lin0, grad_mask,
                                                                     any_frozen)
        loss = float(traj[-1])
        lin = lin_d.detach().cpu()
        values.append(src_val)
        if record is not None:
            record.setdefault("traj", []).extend(traj.tolist())
            record.setdefault("lin", []).append(lin.numpy().copy())
        restarts += 1
        if max_restarts is not None:                                          # retrain_bags: one restart, nothing frozen
            frozen = torch.zeros_like(frozen)
            any_frozen = False
        elif loss < current_loss:
            # retrieve_destinations_low_loss (:530-543) over the dict `arg`: slot k (k-th distinct bag) holds the
            # destination of the LAST bag with that content and is tested against loss_per_bag[k]
            lpb = (diff * diff)[:prob.slot_first.numel()]
            low = lpb < 0.0001
            frozen = torch.zeros_like(frozen)
            frozen[best_dst[prob.slot_last][low]] = True
            any_frozen = bool(low.any())
            current_loss, rest = loss, 0
        else:
            rest += 1
        if max_restarts is None:
            grad_mask[frozen] = 0
        if record is not None:
            record.setdefault("frozen", []).append(frozen.clone())
            record.setdefault("w", []).append(trained_w.clone())
        # reinitialize_weights (:499-516): frozen destinations keep their value, the others are re-drawn in dict order
        redraw = ~frozen[keys]
        k2 = int(redraw.sum())
        new_w = torch.zeros(sg.n, dtype=torch.float32, device=dev)
        new_w[keys[~redraw]] = trained_w[keys[~redraw]]
        if k2:
            new_w[keys[redraw]] = torch.from_numpy(uniform_draws(0.0, 1.0, k2)).to(dev).to(torch.float32)
        weights = new_w
        if max_restarts is not None and restarts >= max_restarts:
            break
    if record is not None:
        record.update(dest_keys=keys, prob=prob)
    vals = torch.stack(values) if values else torch.zeros(0, sg.n, device=dev)
    lin = lin if lin is not None else torch.zeros(int(features_dim))
    return current_loss, lin, vals, prob.visited, prob.skip


def accept_relation(sg, state, rel, lin, values, visited, x_dev, dataset):
    """What the reference does for an accepted relation after retrain_bags (main.py:1432-1435): relabel the visited
    sources (positive iff any restart predicted > 0.9, :596-634), build the new dictionaries over them
    (create_edge_dictionary, BAGS=False, `args.dataset`), drop the sources whose feature . linear weight is < 0.01 and
    take one 0 per dropped edge from their destinations' label lists (clean_dictionaries, :456-477)."""
    dev = sg.device
    best = torch.nan_to_num(values, nan=float("-inf")).amax(dim=0) if values.numel() else torch.full((sg.n,), float("-inf"), device=dev)
    new_labels = torch.zeros(sg.n, dtype=torch.float32, device=dev)
    new_labels[visited] = (best[visited] > 0.9).to(torch.float32)
    rows, cols = sg.rel_edges(rel)
    member = torch.zeros(sg.n, dtype=torch.bool, device=dev)
    member[visited] = True
    in_mask = member[rows]
    rows_m, cols_m = rows[in_mask], cols[in_mask]
    src_lab = _source_labels(sg, rows_m, new_labels, visited.tolist() if dataset != "synthetic" else None, dataset)
    has = torch.zeros(sg.n, dtype=torch.bool, device=dev)
    has[rows] = True
    src_order = visited[has[visited]]
    new = _state_from_edges(sg, rel, src_order, rows_m, cols_m, src_lab, new_labels)
    # clean_dictionaries: the decision a < 0.01 is the reference's torch.dot in float32; values within 1e-6 of the
    # threshold are re-evaluated exactly that way on the host
    lin_d = lin.to(device=dev, dtype=torch.float32)
    a = (x_dev[src_order] * lin_d).sum(dim=1)
    near = torch.nonzero((a - 0.01).abs() < 1e-6).reshape(-1)
    if near.numel():
        xs, lc = x_dev[src_order[near]].cpu(), lin.to(torch.float32).cpu()
        a[near] = torch.stack([torch.dot(xs[i], lc) for i in range(xs.size(0))]).to(dev)
    drop = a < 0.01
    if bool(drop.any()):
        dropped = torch.zeros(sg.n, dtype=torch.bool, device=dev)
        dropped[src_order[drop]] = True
        gone = cols_m[dropped[rows_m]]
        removed = torch.zeros(sg.n, dtype=torch.int32, device=dev).scatter_add_(0, gone, torch.ones(gone.numel(), dtype=torch.int32, device=dev))
        new.count0 = (new.count0 - removed).clamp_(min=0)
        new.src_order = src_order[~drop]
    return new
