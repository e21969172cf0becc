"""B200-native drop-in for the reference's `model.py` MP-GNN model: `MPNetm` keeps its
7-argument constructor, attributes and state_dict keys (model.py:179-228).  Every conv hop
runs the fused CUDA hop (aggregate -> project -> +root +bias -> relu -> dropout); the
fc1/fc2 head runs the library's own dense kernels.
"""
import torch
import torch.nn as nn

from . import _lib
from .graph import RelationGraph, graph_for
from .mp_rgcn_layer import CustomRGCNConv, _workspace


class _LinearFunction(torch.autograd.Function):
    """y = act(x @ W^T + b) with nn.Linear's [out,in] weight, on mpgnn_gemm_rows/gemm_tn."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        lib = _lib.load()
        m, k = x.shape
        n = weight.size(0)
        dev = x.device
        y = torch.empty(m, n, dtype=torch.float32, device=dev)
        ws = _workspace(dev, lib.mpgnn_gemm_workspace_bytes(m, max(k, n), max(k, n)))
        with torch.cuda.device(dev):
            # B(k,n) = W[n,k]  -> ldb_k = 1, ldb_n = k
            rc = lib.mpgnn_gemm_rows(_lib.ptr(x), k, m, k, _lib.ptr(weight), 1, k, n, _lib.ptr(bias), int(relu), None,
                                     0, _lib.ptr(y), n, _lib.ptr(ws), ws.numel(), _lib.current_stream())
        _lib.check(rc)
        ctx.save_for_backward(x, weight, y)
        ctx.relu = bool(relu)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, weight, y = ctx.saved_tensors
        m, k = x.shape
        n = weight.size(0)
        dev = x.device
        gy = gy.contiguous()
        ws = _workspace(dev, lib.mpgnn_gemm_workspace_bytes(m, max(k, n), max(k, n)))
        st = _lib.current_stream
        with torch.cuda.device(dev):
            if ctx.relu:  # g_z = g_y * [y > 0]: identity GEMM with the gate epilogue would waste flops
                gz = torch.where(y > 0, gy, torch.zeros_like(gy))
            else:
                gz = gy
            gx = None
            if ctx.needs_input_grad[0]:
                gx = torch.empty(m, k, dtype=torch.float32, device=dev)
                # g_x = g_z @ W : B(k'=n index, n'=k index) = W[n,k] -> ldb_k = k, ldb_n = 1
                _lib.check(lib.mpgnn_gemm_rows(_lib.ptr(gz), n, m, n, _lib.ptr(weight), k, 1, k, None, 0, None, 0,
                                               _lib.ptr(gx), k, _lib.ptr(ws), ws.numel(), st()))
            gw = torch.empty(n, k, dtype=torch.float32, device=dev)
            gb = torch.empty(n, dtype=torch.float32, device=dev)
            # g_W[n,k] = g_z^T @ x  (deterministic split over the rows); g_b = colsum(g_z) needs B = g_z,
            # so compute g_W^T = x^T @ g_z  [k,n] with the colsum row, then transpose the small result.
            gwt = torch.empty(k, n, dtype=torch.float32, device=dev)
            _lib.check(lib.mpgnn_gemm_tn(_lib.ptr(x), k, m, k, _lib.ptr(gz), n, n, _lib.ptr(gwt), n, _lib.ptr(gb),
                                         _lib.ptr(ws), ws.numel(), st()))
            gw.copy_(gwt.t())
        return gx, gw, gb, None


class MPNetm(torch.nn.Module):
    """model.py:179-228.  One stack of CustomRGCNConv per metapath (layer k consumes relation
    metapaths[i][k]), relu + Dropout(0.6) after every conv, concat -> fc1 -> relu -> fc2 ->
    log_softmax.  `num_rel` and `output_dim` are accepted and ignored, as in the reference.
    Construction order (convs per metapath, then fc1, fc2) matches the reference, so the same
    torch.manual_seed yields the same state_dict."""

    def __init__(self, input_dim, hidden_dim, num_rel, output_dim, ll_output_dim, n_metapaths, metapaths,
                 device=None, precision="tf32x3"):
        super().__init__()
        self.n_metapaths = n_metapaths
        self.metapaths = metapaths
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.precision = precision
        self.layers_list = torch.nn.ModuleList()
        for i in range(len(metapaths)):
            convs = torch.nn.ModuleList()
            convs.append(CustomRGCNConv(input_dim, hidden_dim, 1, flow="target_to_source", device="cpu"))
            for _ in range(len(metapaths[i]) - 1):
                convs.append(CustomRGCNConv(hidden_dim, hidden_dim, 1, flow="target_to_source", device="cpu"))
            self.layers_list.append(convs)
        self.fc1 = torch.nn.Linear(hidden_dim * len(metapaths), hidden_dim)
        self.fc2 = torch.nn.Linear(hidden_dim, ll_output_dim)
        self.log_softmax = torch.nn.LogSoftmax(dim=1)
        self.dropout = nn.Dropout(0.6)
        self.dropout2 = nn.Dropout(0.6)
        self._injected_masks = None
        if device is None:
            device = "cuda" if torch.cuda.is_available() else None
        if device is not None:
            self.to(device)

    def inject_dropout_masks(self, masks):
        """Parity seam: {(metapath, layer): keep-mask [N, hidden]} used instead
        of the counter RNG on the next training-mode forward calls; None restores the RNG."""
        self._injected_masks = masks

    def forward(self, x, edge_index, edge_type=None):
        dev = self.fc1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("MPNetm has no CPU path: move the module to a CUDA device")
        if isinstance(edge_index, RelationGraph):
            graph = edge_index
        else:
            need = 1 + max((int(r) for mp in self.metapaths for r in mp), default=0)
            graph = graph_for(edge_index, edge_type, x.size(0), dev, num_relations=need)
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        embeddings = []
        for i, mp in enumerate(self.metapaths):
            h = x
            for k, rel in enumerate(mp):
                drop = self.dropout if k == 0 else self.dropout2
                p = drop.p if self.training else 0.0
                mask = None
                if self.training and self._injected_masks is not None:
                    mask = self._injected_masks[(i, k)]
                h = self.layers_list[i][k].hop(rel, h, graph, relu=True, dropout_p=p, dropout_mask=mask,
                                               precision=self.precision)
            embeddings.append(h)
        e = embeddings[0] if len(embeddings) == 1 else torch.cat(embeddings, dim=1)
        a1 = _LinearFunction.apply(e, self.fc1.weight, self.fc1.bias, True)
        lg = _LinearFunction.apply(a1, self.fc2.weight, self.fc2.bias, False)
        return self.log_softmax(lg)
