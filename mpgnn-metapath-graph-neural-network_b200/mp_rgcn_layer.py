"""B200-native drop-in for the reference's `mp_rgcn_layer.py`: `masked_edge_index` and
`CustomRGCNConv` keep their names, constructor arguments, attributes, state_dict keys and
forward signature (mp_rgcn_layer.py:29-37, 91-155, 158-271); the work is done by the CUDA
kernels behind the C ABI (include/mpgnn_b200.h).  There is no CPU path.
"""
import math

import torch
from torch import Tensor
from torch.nn import Parameter

from . import _lib
from .graph import RelationGraph, graph_for

_WORKSPACES = {}


def _workspace(device, nbytes):
    """One grow-only scratch buffer per device; all kernels of a call chain run on the
    caller's current stream, so sequential reuse is safe."""
    buf = _WORKSPACES.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _WORKSPACES[device] = buf
    return buf


def masked_edge_index(edge_index, edge_mask):
    """mp_rgcn_layer.py:29-37 -- order-preserving column filter (kept for API parity; the
    layer itself uses the relation-bucketed CSR built once per graph)."""
    if isinstance(edge_index, Tensor):
        return edge_index[:, edge_mask]
    raise NotImplementedError("SparseTensor adjacency is not supported (dead code in the reference)")


def pack_mask_bits(mask):
    """bool/0-1 [N,F] -> uint8 [N, ceil(F/8)], MSB first (numpy.packbits(axis=1) layout)."""
    n, f = mask.shape
    fb = (f + 7) // 8
    m = torch.zeros(n, fb * 8, dtype=torch.uint8, device=mask.device)
    m[:, :f] = (mask != 0).to(torch.uint8)
    w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device=mask.device)
    return (m.view(n, fb, 8) * w).sum(dim=2).to(torch.uint8).contiguous()


class _HopFunction(torch.autograd.Function):
    """y = drop(relu(mean_r(x) @ W + x @ root + bias)) via mpgnn_hop_fwd / mpgnn_hop_bwd."""

    @staticmethod
    def forward(ctx, x, weight, root, bias, graph, relation, flags, dropout_p, seed, offset, mask_bits):
        lib = _lib.load()
        n, f_in = x.shape
        f_out = weight.size(1)
        dev = x.device
        h = torch.empty(n, f_in, dtype=torch.float32, device=dev)
        y = torch.empty(n, f_out, dtype=torch.float32, device=dev)
        # [y > 0] as a bitmask: all the backward needs of y, at 1/32 of its size
        actmask = (torch.empty(n, f_out // 32, dtype=torch.int32, device=dev)
                   if (flags & _lib.F_RELU) and f_out % 32 == 0 else None)
        ws_bytes = lib.mpgnn_hop_workspace_bytes(n, f_in, f_out)
        ws = _workspace(dev, ws_bytes)
        with torch.cuda.device(dev):
            rc = lib.mpgnn_hop_fwd(graph.handle, int(relation), _lib.ptr(x), f_in, _lib.ptr(weight), _lib.ptr(root),
                                   _lib.ptr(bias), f_out, flags, float(dropout_p), int(seed), int(offset),
                                   _lib.ptr(mask_bits), _lib.ptr(h), _lib.ptr(y), _lib.ptr(actmask), _lib.ptr(ws),
                                   ws.numel(), _lib.current_stream())
        _lib.check(rc)
        ctx.use_actmask = actmask is not None
        ctx.save_for_backward(x, h, actmask if ctx.use_actmask else y, weight, root)
        ctx.graph, ctx.relation, ctx.flags, ctx.dropout_p = graph, int(relation), flags, float(dropout_p)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, h, y_or_mask, weight, root = ctx.saved_tensors
        y, actmask = (None, y_or_mask) if ctx.use_actmask else (y_or_mask, None)
        n, f_in = x.shape
        f_out = weight.size(1)
        dev = x.device
        gy = gy.contiguous()
        need_gx = ctx.needs_input_grad[0]
        gx = torch.empty_like(x) if need_gx else None
        gw = torch.empty_like(weight)
        groot = torch.empty_like(root)
        gbias = torch.empty(f_out, dtype=torch.float32, device=dev)
        flags = ctx.flags | (_lib.F_NEED_GX if need_gx else 0)
        ws_bytes = lib.mpgnn_hop_workspace_bytes(n, f_in, f_out)
        ws = _workspace(dev, ws_bytes)
        with torch.cuda.device(dev):
            rc = lib.mpgnn_hop_bwd(ctx.graph.handle, ctx.relation, _lib.ptr(x), _lib.ptr(h), _lib.ptr(y),
                                   _lib.ptr(actmask), _lib.ptr(gy), f_in, _lib.ptr(weight), _lib.ptr(root), f_out, flags,
                                   ctx.dropout_p, _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(groot), _lib.ptr(gbias),
                                   _lib.ptr(ws), ws.numel(), _lib.current_stream())
        _lib.check(rc)
        return gx, gw, groot, (gbias if ctx.has_bias else None), None, None, None, None, None, None, None


class CustomRGCNConv(torch.nn.Module):
    """Single-relation-per-layer RGCN conv of the reference (mp_rgcn_layer.py:40-283):

        out = mean_{j in N_r(i)} x_j @ weight + x_i @ root + bias

    with ONE [in,out] weight, aggregate-then-project, relation chosen per call.
    Constructor arguments and attributes mirror the reference.  Parameters are drawn on the
    CPU generator in the reference's order (weight, root glorot-uniform; bias zeros) and then
    moved to `device`, so the same `torch.manual_seed` gives the same state_dict.

    Unsupported (dead in the reference): num_bases / num_blocks, integer
    `x`, tuple `x`, SparseTensor adjacency, aggr other than 'mean', flow other than
    'target_to_source' -- each raises NotImplementedError.
    """

    def __init__(self, in_channels, out_channels, num_relations, num_bases=None, num_blocks=None, aggr="mean",
                 root_weight=True, bias=True, device=None, precision="tf32x3", **kwargs):
        super().__init__()
        self.precision = precision
        kwargs.setdefault("aggr", aggr)
        if num_bases is not None and num_blocks is not None:
            raise ValueError("Can not apply both basis-decomposition and "
                             "block-diagonal-decomposition at the same time.")
        if num_bases is not None or num_blocks is not None:
            raise NotImplementedError("basis/block decomposition is dead code in the reference and not built")
        self.aggr = kwargs.pop("aggr")
        self.flow = kwargs.pop("flow", "source_to_target")
        self.node_dim = kwargs.pop("node_dim", 0)
        if kwargs:
            raise TypeError("unexpected arguments: %s" % sorted(kwargs))
        if self.aggr != "mean":
            raise NotImplementedError("only aggr='mean' (the reference default) is built")
        if self.flow != "target_to_source":
            raise NotImplementedError("only flow='target_to_source' (what model.py:190-192 passes) is built")
        if isinstance(in_channels, (tuple, list)):
            raise NotImplementedError("bipartite in_channels are dead code in the reference and not built")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_relations = num_relations
        self.num_bases = num_bases
        self.num_blocks = num_blocks
        self.in_channels_l = in_channels
        self.weight = Parameter(torch.empty(in_channels, out_channels))
        self.register_parameter("comp", None)
        if root_weight:
            self.root = Parameter(torch.empty(in_channels, out_channels))
        else:
            self.register_parameter("root", None)
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        if device is None:
            device = "cuda" if torch.cuda.is_available() else None
        if device is not None:
            self.to(device)

    def reset_parameters(self):
        """glorot(weight); glorot(root); zeros(bias) -- mp_rgcn_layer.py:151-155."""
        with torch.no_grad():
            for t in (self.weight, self.root):
                if t is not None:
                    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
                    if t.is_cuda:
                        t.copy_(torch.empty(t.shape).uniform_(-a, a))  # keep the CPU RNG stream
                    else:
                        t.uniform_(-a, a)
            if self.bias is not None:
                self.bias.zero_()

    # -- internal: one hop with the epilogue MPNetm wants fused ---------------------------
    def hop(self, relation, x, graph, relu=False, dropout_p=0.0, dropout_mask=None, seed=None, offset=0,
            precision=None, h_layout="auto"):
        """`precision`: "tf32x3" (default, `self.precision`) = projection and weight gradient on tcgen05 with the
        error-compensated 3xTF32 split (fp32-class, <= 2e-6 measured) wherever the shape is eligible, the exact-fp32
        SIMT kernels otherwise; "fp32" = the SIMT kernels always."""
        precision = self.precision if precision is None else precision
        # h_layout: "auto" = compact (one row of aggregated features per node with edges of the relation) on large sparse
        # relations, dense otherwise; "compact" / "dense" force it where the shapes allow (tests, benchmarks)
        if not self.weight.is_cuda:
            raise RuntimeError("CustomRGCNConv has no CPU path: move the module to a CUDA device")
        if self.root is None:
            raise NotImplementedError("root_weight=False is not built (the reference never uses it)")
        if isinstance(x, tuple) or x is None or not torch.is_floating_point(x):
            raise NotImplementedError("only a float feature matrix x is supported (reference live branch)")
        dev = self.weight.device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        flags = _lib.F_RELU if relu else 0
        mask_bits = None
        if dropout_mask is not None:
            flags |= _lib.F_DROPOUT_MASK
            mask_bits = pack_mask_bits(dropout_mask.to(dev))
        elif dropout_p > 0.0:
            flags |= _lib.F_DROPOUT_SEED
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if precision == "tf32x3":
            flags |= _lib.F_TF32X3
        elif precision == "bf16":
            raise NotImplementedError("precision='bf16' is not built (MPGNN_F_BF16 is rejected by the library)")
        elif precision != "fp32":
            raise ValueError("precision must be 'fp32' or 'tf32x3'")
        if h_layout not in ("auto", "compact", "dense"):
            raise ValueError("h_layout must be 'auto', 'compact' or 'dense'")
        flags |= {"auto": 0, "compact": _lib.F_COMPACT_H, "dense": _lib.F_DENSE_H}[h_layout]
        if (flags & (_lib.F_DROPOUT_MASK | _lib.F_DROPOUT_SEED)) and not relu:
            raise NotImplementedError("dropout without relu is not built (MPNetm always applies relu first, "
                                      "model.py:210-214)")
        return _HopFunction.apply(x, self.weight, self.root, self.bias, graph, int(relation), flags,
                                  float(dropout_p), seed or 0, offset, mask_bits)

    def forward(self, layer_num, relation, x, edge_index, edge_type=None):
        """Reference signature (mp_rgcn_layer.py:158-159).  `layer_num` is unused, as in the
        reference.  `edge_index` may also be a prebuilt RelationGraph."""
        if isinstance(edge_index, RelationGraph):
            graph = edge_index
        else:
            if not isinstance(edge_index, Tensor):
                raise NotImplementedError("SparseTensor adjacency is not supported")
            assert edge_type is not None
            n = x.size(0)
            # a relation id the edge list never uses is an empty neighbourhood, not an error (:231)
            graph = graph_for(edge_index, edge_type, n, self.weight.device, num_relations=int(relation) + 1)
        return self.hop(relation, x, graph)

    def message(self, x_j):
        return x_j

    def __repr__(self):
        return "%s(%s, %s, num_relations=%s)" % (self.__class__.__name__, self.in_channels, self.out_channels,
                                                 self.num_relations)
