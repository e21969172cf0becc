"""The reference's all-relation RGCN comparison model on the same kernels: `Net` (model.py:132-151) and
the training call of main_rgcn.py:452-472.  `Net` stacks two `torch_geometric.nn.RGCNConv(in, out, num_relations,
flow='target_to_source')`; that class is third-party (torch-geometric==2.3.1, requirements.txt:7, not vendored), its
published algorithm restated here:

    out = sum_r  mean_{j in N_r(i)} x_j @ weight[r]  +  x_i @ root  +  bias        (aggr='mean', per-relation weights)

i.e. the MP-RGCN hop summed over ALL relations instead of one per layer.  On the device it is a loop over the relation
buckets of the graph handle (K1) with the per-hop aggregation kernel (K2: `mpgnn_spmm`) writing each relation's mean
into its column block of H = [h_0 | ... | h_{R-1} | x], ONE projection H @ [W_0; ...; W_{R-1}; root] + bias
(`mpgnn_gemm_rows`), and the matching backward: g_[W;root] = H^T g (`mpgnn_gemm_tn`, deterministic split),
g_H = g [W;root]^T, g_x = g_H[:, x block] + sum_r A_r^T D_r^-1 g_H[:, block r] (`mpgnn_scale_rows_by_degree` +
transposed `mpgnn_spmm`, in place).  No CPU path.
"""
import math

import torch
from torch.nn import Parameter

from . import _lib
from .graph import RelationGraph, graph_for
from .mp_rgcn_layer import _workspace
from .model import _LinearFunction

H_BYTES_LIMIT = 16 << 30      # [N, (R+1) F_in] is materialised: meant for the comparison runs at C1-C3 sizes


class _RGCNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, root, bias, graph, relu):
        lib = _lib.load()
        n, f_in = x.shape
        r, _, f_out = weight.shape
        dev = x.device
        k = (r + 1) * f_in
        if n * k * 4 > H_BYTES_LIMIT:
            raise NotImplementedError("RGCNConv baseline: [N, (R+1) F_in] = %.1f GB exceeds the materialisation limit"
                                      % (n * k * 4 / 2 ** 30))
        st = _lib.current_stream
        r_graph = min(r, graph.num_relations)        # a relation the edge list never uses is an empty neighbourhood
        with torch.cuda.device(dev):
            hcat = (torch.empty if r_graph == r else torch.zeros)(n, k, dtype=torch.float32, device=dev)
            for rel in range(r_graph):               # K2 per relation, straight into its column block
                _lib.check(lib.mpgnn_spmm(graph.handle, rel, 0, 1, _lib.ptr(x), f_in, f_in, None, 0,
                                          ctypes_offset(hcat, rel * f_in), k, st()))
            hcat[:, r * f_in:] = x
            wcat = torch.cat([weight.reshape(r * f_in, f_out), root], dim=0).contiguous()          # [K, f_out]
            y = torch.empty(n, f_out, dtype=torch.float32, device=dev)
            ws = _workspace(dev, _ws_bytes(lib, n, k, f_out))
            _lib.check(lib.mpgnn_gemm_rows(_lib.ptr(hcat), k, n, k, _lib.ptr(wcat), f_out, 1, f_out, _lib.ptr(bias), int(relu),
                                           None, 0, _lib.ptr(y), f_out, _lib.ptr(ws), ws.numel(), st()))
        ctx.save_for_backward(hcat, wcat, y)
        ctx.graph, ctx.relu, ctx.dims = graph, bool(relu), (n, f_in, f_out, r)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        hcat, wcat, y = ctx.saved_tensors
        n, f_in, f_out, r = ctx.dims
        k = (r + 1) * f_in
        dev = gy.device
        gy = gy.contiguous()
        st = _lib.current_stream
        with torch.cuda.device(dev):
            gz = torch.where(y > 0, gy, torch.zeros_like(gy)) if ctx.relu else gy
            ws = _workspace(dev, _ws_bytes(lib, n, k, f_out))
            gwcat = torch.empty(k, f_out, dtype=torch.float32, device=dev)
            gb = torch.empty(f_out, dtype=torch.float32, device=dev)
            _lib.check(lib.mpgnn_gemm_tn(_lib.ptr(hcat), k, n, k, _lib.ptr(gz), f_out, f_out, _lib.ptr(gwcat), f_out,
                                         _lib.ptr(gb), _lib.ptr(ws), ws.numel(), st()))
            gx = None
            if ctx.needs_input_grad[0]:
                gh = torch.empty(n, k, dtype=torch.float32, device=dev)
                # g_H = g_z @ wcat^T : B(k'=o, n'=j) = wcat[j, o] -> ldb_k = 1, ldb_n = f_out
                _lib.check(lib.mpgnn_gemm_rows(_lib.ptr(gz), f_out, n, f_out, _lib.ptr(wcat), 1, f_out, k, None, 0, None, 0,
                                               _lib.ptr(gh), k, _lib.ptr(ws), ws.numel(), st()))
                gx = gh[:, r * f_in:].contiguous()
                for rel in range(min(r, ctx.graph.num_relations)):
                    blk = ctypes_offset(gh, rel * f_in)
                    _lib.check(lib.mpgnn_scale_rows_by_degree(ctx.graph.handle, rel, blk, k, f_in, blk, k, st()))
                    _lib.check(lib.mpgnn_spmm(ctx.graph.handle, rel, 1, 0, blk, k, f_in, _lib.ptr(gx), f_in, _lib.ptr(gx),
                                              f_in, st()))
        gw = gwcat[:r * f_in].reshape(r, f_in, f_out)
        groot = gwcat[r * f_in:]
        return gx, gw, groot, (gb if ctx.has_bias else None), None, None


def _ws_bytes(lib, n, k, f_out):
    """Scratch for the three dense calls of a layer: H @ Wcat, H^T g, g @ Wcat^T."""
    return max(lib.mpgnn_gemm_workspace_bytes(n, k, f_out), lib.mpgnn_gemm_workspace_bytes(n, f_out, k))


def ctypes_offset(t, col):
    """Device pointer to column `col` of row 0 of a row-major float32 matrix."""
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + 4 * int(col))


class RGCNConv(torch.nn.Module):
    """torch_geometric.nn.RGCNConv (2.3.1) as model.py:137-138 builds it: per-relation weights [R, in, out], root
    [in, out], bias [out]; glorot(weight), glorot(root), zeros(bias) in that order (same torch.manual_seed => same
    state_dict); aggr='mean', flow='target_to_source'.  Basis / block decomposition are not built."""

    def __init__(self, in_channels, out_channels, num_relations, num_bases=None, num_blocks=None, aggr="mean",
                 root_weight=True, bias=True, device=None, **kwargs):
        super().__init__()
        if num_bases is not None or num_blocks is not None:
            raise NotImplementedError("basis/block decomposition is not built (the reference never passes it)")
        if aggr != "mean" or kwargs.pop("flow", "source_to_target") != "target_to_source" or not root_weight:
            raise NotImplementedError("only aggr='mean', flow='target_to_source', root_weight=True (model.py:137-138)")
        self.in_channels, self.out_channels, self.num_relations = in_channels, out_channels, num_relations
        self.weight = Parameter(torch.empty(num_relations, in_channels, out_channels))
        self.root = Parameter(torch.empty(in_channels, out_channels))
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
        with torch.no_grad():
            for t in (self.weight, self.root):
                a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
                t.copy_(torch.empty(t.shape).uniform_(-a, a))         # CPU generator, like the reference
            if self.bias is not None:
                self.bias.zero_()

    def forward(self, x, edge_index, edge_type=None, relu=False):
        if not self.weight.is_cuda:
            raise RuntimeError("RGCNConv has no CPU path: move the module to a CUDA device")
        dev = self.weight.device
        graph = edge_index if isinstance(edge_index, RelationGraph) else graph_for(
            edge_index, edge_type, x.size(0), dev, num_relations=self.num_relations)
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        return _RGCNFunction.apply(x, self.weight, self.root, self.bias, graph, relu)


class Net(torch.nn.Module):
    """model.py:132-151: conv1 (input -> hidden), conv2 (hidden -> output) applied `metapath_length - 1` times, relu after
    every conv, LinearLayer, log_softmax."""

    def __init__(self, input_dim, hidden_dim, num_rel, output_dim, ll_output_dim, metapath_length, device=None):
        super().__init__()
        self.metapath_length = metapath_length
        self.conv1 = RGCNConv(input_dim, hidden_dim, num_rel, flow="target_to_source", device="cpu")
        self.conv2 = RGCNConv(hidden_dim, output_dim, num_rel, flow="target_to_source", device="cpu")
        self.LinearLayer = torch.nn.Linear(output_dim, ll_output_dim)
        if device is None:
            device = "cuda" if torch.cuda.is_available() else None
        if device is not None:
            self.to(device)

    def forward(self, x, edge_index, edge_type=None):
        dev = self.LinearLayer.weight.device
        graph = edge_index if isinstance(edge_index, RelationGraph) else graph_for(
            edge_index, edge_type, x.size(0), dev, num_relations=self.conv1.num_relations)
        for layer_index in range(self.metapath_length):
            x = (self.conv1 if layer_index == 0 else self.conv2)(x, graph, relu=True)      # F.relu fused into the GEMM epilogue
        x = _LinearFunction.apply(x, self.LinearLayer.weight, self.LinearLayer.bias, False)
        return torch.log_softmax(x, dim=1)


def mpgnn_parallel_multiple(data_mpgnn, input_dim, hidden_dim, num_rel, output_dim, ll_output_dim, metapath_length,
                            epochs=999, log=print):
    """main_rgcn.py:452-472: 999 x (train, validation, test) of `Net`, then the test macro-F1 of `best_model` -- an
    alias of the model, not a copy, so it is the final model's.  Returns that number."""
    from .main import mpgnn_train, mpgnn_validation, mpgnn_test, ADAM_LR, ADAM_WEIGHT_DECAY
    model = Net(input_dim, hidden_dim, num_rel, output_dim, ll_output_dim, metapath_length)
    optimizer = torch.optim.Adam(model.parameters(), lr=ADAM_LR, weight_decay=ADAM_WEIGHT_DECAY)
    class_weight = None
    for epoch in range(1, epochs + 1):
        loss, class_weight = mpgnn_train(model, optimizer, data_mpgnn)
        train_acc, f1_val_micro, _, loss_val = mpgnn_validation(model, data_mpgnn, class_weight)
        if epoch % 10 == 0 and log is not None:
            _, f1_micro_test = mpgnn_test(model, data_mpgnn, class_weight)
            log(epoch, "train loss %0.3f" % loss, "validation loss %0.3f" % loss_val, "train micro: %0.3f" % train_acc,
                "validation micro: %0.3f" % f1_val_micro, "test micro: %0.3f" % f1_micro_test)
    test_loss, f1_micro_test = mpgnn_test(model, data_mpgnn, class_weight)
    if log is not None:
        log("test loss %0.3f" % test_loss, "test micro %0.3f" % f1_micro_test)
    return f1_micro_test


rgcn_parallel_multiple = mpgnn_parallel_multiple
