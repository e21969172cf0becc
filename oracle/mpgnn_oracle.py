"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch-CPU fp32 + numpy integers) of the
MPGNN hot path.  It is the checker for the CUDA path, never the thing shipped:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.

Parity status: the reference has no tests.  Integer-side results are
PINNED by the reference's own fixtures (edges.pkl / labels.pkl / node_features.pkl,
see tests/test_oracle_golden.py); floating-point results are pinned against outputs of
the UNMODIFIED reference sources run in the build container behind dependency
stand-ins (oracle/ref_shims.py -> tests/golden/*.npz, generator
tests/golden/make_golden.py).  The message-passing arithmetic itself lives in
torch_geometric==2.3.1 (requirements.txt:7), which is not vendored in /root/reference;
its published scatter-mean algorithm is restated in `propagate_mean` below.

Every function cites the reference lines it follows (paths relative to /root/reference).
Orientation: edge_index[0] is the aggregation TARGET (CSR row),
edge_index[1] the message SOURCE (CSR column), because every conv is built with
flow='target_to_source' (model.py:190,192).
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# a1: relation filter and the stable (relation,row) / (relation,col) bucketing
# --------------------------------------------------------------------------------------


def masked_edge_index(edge_index, edge_mask):
    """mp_rgcn_layer.py:29-37 -- order-preserving column filter."""
    return edge_index[:, edge_mask]


def relation_csr(edge_index, edge_type, num_nodes, num_relations, transpose=False):
    """Stable bucketing of the edge list by (relation, row) -- what repeated calls of
    `masked_edge_index(edge_index, edge_type == r)` (mp_rgcn_layer.py:231) followed by
    PyG's scatter at edge_index[0] enumerate.  Inside one (relation,row) bucket the
    original edge order is kept, duplicates included (edges.pkl
    counts duplicate triplets twice).

    Returns (ptr int64 [R*N+1], other int32 [E], perm int32 [E]) where bucket
    (r, i) owns positions ptr[r*N+i] .. ptr[r*N+i+1] and `other` holds the column
    (message source) of each edge -- or, with transpose=True, buckets are keyed by
    (relation, col) and `other` holds the row.  `perm` is the original edge id.
    """
    ei = np.asarray(edge_index)
    et = np.asarray(edge_type).astype(np.int64)
    key_nodes = ei[1] if transpose else ei[0]
    other = ei[0] if transpose else ei[1]
    key = et * np.int64(num_nodes) + key_nodes.astype(np.int64)
    perm = np.argsort(key, kind="stable")
    counts = np.bincount(key, minlength=num_relations * num_nodes)
    ptr = np.zeros(num_relations * num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    return ptr, other[perm].astype(np.int32), perm.astype(np.int32)


# --------------------------------------------------------------------------------------
# a4: PyG 2.3.1 MessagePassing.propagate(aggr='mean', flow='target_to_source')
# --------------------------------------------------------------------------------------


def propagate_mean(x, tmp):
    """mp_rgcn_layer.py:236 + :274-275 -> torch_geometric 2.3.1 propagate / scatter-mean:
    gather x at tmp[1], scatter_add_ at tmp[0] in edge order, divide by the count
    clamped to >= 1 (rows with no edge stay 0)."""
    n = x.size(0)
    x_j = x.index_select(0, tmp[1])
    out = x.new_zeros(n, x.size(1))
    out.scatter_add_(0, tmp[0].view(-1, 1).expand_as(x_j), x_j)
    cnt = x.new_zeros(n).scatter_add_(0, tmp[0], x.new_ones(tmp.size(1))).clamp(min=1)
    return out / cnt.view(-1, 1), cnt


# --------------------------------------------------------------------------------------
# a2/a3: CustomRGCNConv
# --------------------------------------------------------------------------------------


def glorot_(t):
    """torch_geometric.nn.inits.glorot (mp_rgcn_layer.py:152-154)."""
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    t.data.uniform_(-a, a)
    return t


def conv_init(in_channels, out_channels):
    """CustomRGCNConv.__init__/reset_parameters (mp_rgcn_layer.py:91-155): weight then
    root drawn glorot-uniform from the global torch CPU generator, bias zeros."""
    w = glorot_(torch.empty(in_channels, out_channels))
    r = glorot_(torch.empty(in_channels, out_channels))
    b = torch.zeros(out_channels)
    return {"weight": w, "root": r, "bias": b}


def conv_forward(x, edge_index, edge_type, relation, weight, root, bias):
    """CustomRGCNConv.forward (mp_rgcn_layer.py:158-271) for the only live branch
    (var_bool=True, float x): out = mean_r(x) @ weight + x @ root + bias.
    Returns (out, h, cnt)."""
    tmp = masked_edge_index(edge_index, edge_type == int(relation))  # :231
    h, cnt = propagate_mean(x, tmp)  # :236
    out = torch.zeros(x.size(0), weight.size(1)) + (h @ weight)  # :198, :245
    out = out + x @ root  # :265
    out = out + bias  # :268
    return out, h, cnt


def conv_backward(x, edge_index, edge_type, relation, weight, root, h, cnt, g_out, need_gx=True):
    """What autograd derives for conv_forward:
    g_bias = sum_i g, g_root = x^T g, g_W = h^T g, t = (g W^T)/cnt,
    g_x[j] = (g root^T)[j] + sum_{e in E_r, col(e)=j} t[row(e)]."""
    tmp = masked_edge_index(edge_index, edge_type == int(relation))
    g_bias = g_out.sum(0)
    g_root = x.t() @ g_out
    g_w = h.t() @ g_out
    g_x = None
    if need_gx:
        t = (g_out @ weight.t()) / cnt.view(-1, 1)
        g_x = g_out @ root.t()
        g_x.index_add_(0, tmp[1], t.index_select(0, tmp[0]))
    return g_x, g_w, g_root, g_bias


# --------------------------------------------------------------------------------------
# a5/a6: MPNetm
# --------------------------------------------------------------------------------------

DROPOUT_P = 0.6  # model.py:200-201


def mpnetm_init(input_dim, hidden_dim, ll_output_dim, metapaths):
    """MPNetm.__init__ (model.py:180-201): same construction order, so the same
    torch.manual_seed gives the same state_dict as the reference."""
    sd = {}
    for i, mp in enumerate(metapaths):
        for k in range(len(mp)):
            p = conv_init(input_dim if k == 0 else hidden_dim, hidden_dim)
            for name, v in p.items():
                sd["layers_list.%d.%d.%s" % (i, k, name)] = v
    fc1 = torch.nn.Linear(hidden_dim * len(metapaths), hidden_dim)
    fc2 = torch.nn.Linear(hidden_dim, ll_output_dim)
    sd["fc1.weight"], sd["fc1.bias"] = fc1.weight.data, fc1.bias.data
    sd["fc2.weight"], sd["fc2.bias"] = fc2.weight.data, fc2.bias.data
    return sd


def mpnetm_forward(sd, x, edge_index, edge_type, metapaths, masks=None, keep=None):
    """MPNetm.forward (model.py:203-228).  `masks[(i,k)]` is the Bernoulli keep-mask of
    the dropout after conv (i,k) in train mode (None = eval mode).  `keep`, when a dict,
    receives the intermediates the explicit backward needs."""
    embs = []
    scale = 1.0 / (1.0 - DROPOUT_P)
    for i, mp in enumerate(metapaths):
        h_in = x
        for k, rel in enumerate(mp):
            pre = "layers_list.%d.%d." % (i, k)
            z, h_agg, cnt = conv_forward(h_in, edge_index, edge_type, rel,
                                         sd[pre + "weight"], sd[pre + "root"], sd[pre + "bias"])
            y = torch.relu(z)  # model.py:210/213
            if masks is not None:
                y = y * masks[(i, k)] * scale  # model.py:211/214 (nn.Dropout(0.6), train)
            if keep is not None:
                keep[(i, k)] = (h_in, h_agg, cnt, y)
            h_in = y
        embs.append(h_in)
    e = torch.cat(embs, dim=1)  # model.py:220
    a1 = torch.relu(e @ sd["fc1.weight"].t() + sd["fc1.bias"])  # :223
    lg = a1 @ sd["fc2.weight"].t() + sd["fc2.bias"]  # :225
    logp = torch.log_softmax(lg, dim=1)  # :226
    if keep is not None:
        keep["e"], keep["a1"], keep["logp"] = e, a1, logp
    return logp


def nll_loss_on_index(logp, idx, y):
    """F.nll_loss(out[idx].squeeze(-1), y) (main.py:1065, 1088, 1106), mean reduction."""
    idx = torch.as_tensor(idx, dtype=torch.long)
    y = torch.as_tensor(y, dtype=torch.long)
    return -(logp[idx, y]).mean()


def mpnetm_loss_and_grads(sd, x, edge_index, edge_type, metapaths, train_idx, train_y, masks=None):
    """One mpgnn_train forward/backward (main.py:1055-1078) with the backward written
    out by hand -- the algorithm the CUDA path implements.
    Returns (loss, grads dict keyed like the state_dict, logp)."""
    keep = {}
    logp = mpnetm_forward(sd, x, edge_index, edge_type, metapaths, masks, keep)
    idx = torch.as_tensor(train_idx, dtype=torch.long)
    y = torch.as_tensor(train_y, dtype=torch.long)
    loss = -(logp[idx, y]).mean()
    n_tr = idx.numel()
    g = {}
    # d loss / d logits = (softmax - onehot)/n on train rows (duplicates in idx accumulate)
    g_lg = torch.zeros_like(logp)
    row_g = torch.exp(logp[idx])
    row_g[torch.arange(n_tr), y] -= 1.0
    g_lg.index_add_(0, idx, row_g / n_tr)
    g["fc2.weight"] = g_lg.t() @ keep["a1"]
    g["fc2.bias"] = g_lg.sum(0)
    g_a1 = (g_lg @ sd["fc2.weight"]) * (keep["a1"] > 0)
    g["fc1.weight"] = g_a1.t() @ keep["e"]
    g["fc1.bias"] = g_a1.sum(0)
    g_e = g_a1 @ sd["fc1.weight"]
    hd = sd["fc1.weight"].size(0)
    scale = 1.0 / (1.0 - DROPOUT_P)
    for i, mp in enumerate(metapaths):
        g_y = g_e[:, i * hd:(i + 1) * hd]
        for k in reversed(range(len(mp))):
            pre = "layers_list.%d.%d." % (i, k)
            h_in, h_agg, cnt, y_out = keep[(i, k)]
            g_z = g_y * (y_out > 0)
            if masks is not None:
                g_z = g_z * scale
            g_x, g_w, g_r, g_b = conv_backward(h_in, edge_index, edge_type, mp[k], sd[pre + "weight"],
                                               sd[pre + "root"], h_agg, cnt, g_z, need_gx=(k > 0))
            g[pre + "weight"], g[pre + "root"], g[pre + "bias"] = g_w, g_r, g_b
            g_y = g_x
    return loss, g, logp


# --------------------------------------------------------------------------------------
# a7: optimiser (torch.optim.Adam(lr=0.01, weight_decay=5e-4), main.py:1119)
# --------------------------------------------------------------------------------------


def adam_step(sd, grads, state, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=5e-4):
    """torch.optim.Adam single-tensor update (classic L2: wd added to the gradient)."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    bc1 = 1.0 - beta1 ** t
    bc2 = 1.0 - beta2 ** t
    for k in sd:
        gk = grads[k]
        if weight_decay != 0.0:
            gk = gk + weight_decay * sd[k]
        m = state.setdefault("m." + k, torch.zeros_like(sd[k]))
        v = state.setdefault("v." + k, torch.zeros_like(sd[k]))
        m.lerp_(gk, 1.0 - beta1)
        v.mul_(beta2).addcmul_(gk, gk, value=1.0 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        sd[k] = sd[k] - (lr / bc1) * (m / denom)
    return sd


# --------------------------------------------------------------------------------------
# a8: macro-F1 (sklearn.metrics.f1_score(average='macro'), main.py:1095-1099, 1112)
# --------------------------------------------------------------------------------------


def macro_f1(pred, true):
    """sklearn f1_score(pred, true, average='macro'): labels = sorted union of both
    arrays; per-label F1 = 2tp/(2tp+fp+fn) with 0 when the denominator is 0."""
    pred = np.asarray(pred).astype(np.int64).ravel()
    true = np.asarray(true).astype(np.int64).ravel()
    labels = np.union1d(pred, true)
    f = []
    for c in labels:
        tp = np.sum((pred == c) & (true == c))
        fp = np.sum((pred == c) & (true != c))
        fn = np.sum((pred != c) & (true == c))
        d = 2 * tp + fp + fn
        f.append(0.0 if d == 0 else 2.0 * tp / d)
    return float(np.mean(f)) if len(f) else 0.0


# --------------------------------------------------------------------------------------
# a9: one candidate scored (mpgnn_parallel_multiple, main.py:1117-1134)
# --------------------------------------------------------------------------------------


def score_candidate(sd, data, metapaths, epochs=999, mask_fn=None, return_trace=False):
    """`epochs` x (mpgnn_train, mpgnn_validation); returns the LAST epoch's validation
    macro-F1 (main.py:1121-1134).  `mask_fn(epoch, (i,k), shape)` supplies dropout keep
    masks (None = dropout disabled -- the deterministic parity variant)."""
    sd = {k: v.clone() for k, v in sd.items()}
    state = {}
    trace = []
    f1_val = 0.0
    for ep in range(1, epochs + 1):
        masks = None
        if mask_fn is not None:
            masks = {}
            for i, mp in enumerate(metapaths):
                for k in range(len(mp)):
                    masks[(i, k)] = mask_fn(ep, (i, k), (data["x"].size(0), sd["fc1.weight"].size(0)))
        loss, grads, _ = mpnetm_loss_and_grads(sd, data["x"], data["edge_index"], data["edge_type"], metapaths,
                                               data["train_idx"], data["train_y"], masks)
        sd = adam_step(sd, grads, state)
        logp = mpnetm_forward(sd, data["x"], data["edge_index"], data["edge_type"], metapaths)
        pred = logp.argmax(1)
        f1_tr = macro_f1(pred[torch.as_tensor(data["train_idx"])], data["train_y"])
        f1_val = macro_f1(pred[torch.as_tensor(data["val_idx"])], data["val_y"])
        loss_val = float(nll_loss_on_index(logp, data["val_idx"], data["val_y"]))
        if return_trace:
            trace.append((float(loss), loss_val, f1_tr, f1_val))
    if return_trace:
        return f1_val, sd, trace
    return f1_val


# --------------------------------------------------------------------------------------
# f4: the all-relation comparison model (model.py:132-151 `Net`; main_rgcn.py:452-472)
# --------------------------------------------------------------------------------------


def rgcn_conv_init(in_channels, out_channels, num_relations):
    """torch_geometric.nn.RGCNConv.reset_parameters (PyG 2.3.1, third-party): glorot(weight [R, in, out]), glorot(root),
    zeros(bias), in that order from the global torch CPU generator."""
    w = glorot_(torch.empty(num_relations, in_channels, out_channels))
    r = glorot_(torch.empty(in_channels, out_channels))
    return {"weight": w, "root": r, "bias": torch.zeros(out_channels)}


def rgcn_conv_forward(x, edge_index, edge_type, weight, root, bias):
    """RGCNConv.forward without decomposition / pyg_lib: the per-relation loop the reference's CustomRGCNConv was cut
    down from (mp_rgcn_layer.py:249-258): out = sum_r mean_r(x) @ weight[r] + x @ root + bias."""
    out = torch.zeros(x.size(0), weight.size(2))
    for r in range(weight.size(0)):
        h, _ = propagate_mean(x, masked_edge_index(edge_index, edge_type == r))
        out = out + h @ weight[r]
    return out + x @ root + bias


def net_init(input_dim, hidden_dim, num_rel, output_dim, ll_output_dim):
    """Net.__init__ (model.py:133-139): conv1, conv2, LinearLayer in construction order; state_dict keys of the model."""
    c1 = rgcn_conv_init(input_dim, hidden_dim, num_rel)
    c2 = rgcn_conv_init(hidden_dim, output_dim, num_rel)
    lin = torch.nn.Linear(output_dim, ll_output_dim)
    sd = {"conv1." + k: v for k, v in c1.items()}
    sd.update({"conv2." + k: v for k, v in c2.items()})
    sd.update({"LinearLayer.weight": lin.weight.detach().clone(), "LinearLayer.bias": lin.bias.detach().clone()})
    return sd


def net_forward(sd, x, edge_index, edge_type, metapath_length):
    """Net.forward (model.py:141-150): relu(conv1), then relu(conv2) metapath_length - 1 times, LinearLayer, log_softmax."""
    for layer in range(metapath_length):
        p = "conv1." if layer == 0 else "conv2."
        x = torch.relu(rgcn_conv_forward(x, edge_index, edge_type, sd[p + "weight"], sd[p + "root"], sd[p + "bias"]))
    return torch.log_softmax(x @ sd["LinearLayer.weight"].t() + sd["LinearLayer.bias"], dim=1)
