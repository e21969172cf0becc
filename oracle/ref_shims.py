"""TEST INFRASTRUCTURE ONLY -- dependency stand-ins that let the UNMODIFIED reference
sources under /root/reference be imported in this container.

Users: `tests/golden/make_golden*.py`, the `-m "not gpu"` validation tests (which skip when
/root/reference is absent) and `bench.py`'s CPU legs (`--impl reference`, `cpu_baseline`), which
run the UNMODIFIED sources vendored into `oracle/_ref/` by `oracle/build_ref.py` as the timed
reference arm.  Nothing on the product path or in the `-m gpu` tests imports it.

The reference needs torch_geometric 2.3.1 / torch_scatter / torch_sparse / mpi4py /
seaborn / mlxtend / imblearn / matplotlib (reference requirements.txt:1-8 and
main.py:1-28, mp_rgcn_layer.py:8-14); none exist here and there is no network.  The
stand-ins restate the published behaviour of exactly the names the reference uses
(SURVEY.md Appendix A):

* torch_geometric.nn.conv.MessagePassing.propagate  (PyG 2.3.1, aggr='mean'):
  x_j = x.index_select(0, edge_index[j]); scatter-sum at edge_index[i]; divide by the
  per-target count clamped to >= 1.  (i, j) = (0, 1) for flow='target_to_source'.
* torch_geometric.nn.inits.glorot / zeros.
* torch_geometric.data.Data: attribute bag, clone() deep, copy.copy shallow.
* mpi4py.MPI.COMM_WORLD: single rank, bcast/gather identity.
"""
import copy
import math
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MPGNN_REFERENCE_ROOT", "/root/reference")
VENDORED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/build_ref.py


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mp_rgcn_layer.py"))


def use_vendored():
    """Point the stand-ins at oracle/_ref (the unmodified hot-path sources copied there by oracle/build_ref.py) -- the
    form in which the reference reaches the GPU box.  Returns False when the folder is absent."""
    global REFERENCE_ROOT
    if not os.path.isfile(os.path.join(VENDORED_ROOT, "mp_rgcn_layer.py")):
        return False
    REFERENCE_ROOT = VENDORED_ROOT
    return True


class _Data:
    """torch_geometric.data.Data stand-in (reference main.py:1245, 1274)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def clone(self):
        return copy.deepcopy(self)

    def __copy__(self):
        new = _Data.__new__(_Data)
        new.__dict__.update(self.__dict__)
        return new


class _MessagePassing(torch.nn.Module):
    """torch_geometric.nn.conv.MessagePassing stand-in, PyG 2.3.1 semantics for the
    one call the reference makes (mp_rgcn_layer.py:236)."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs["x"]
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        n_out = size[i] if size is not None and size[i] is not None else x.size(0)
        x_j = x.index_select(0, edge_index[j])
        msg = self.message(x_j)
        idx = edge_index[i]
        out = msg.new_zeros((n_out,) + tuple(msg.shape[1:]))
        out.scatter_add_(0, idx.view(-1, *([1] * (msg.dim() - 1))).expand_as(msg), msg)
        if self.aggr == "mean":
            cnt = msg.new_zeros(n_out).scatter_add_(0, idx, msg.new_ones(idx.numel()))
            cnt = cnt.clamp(min=1)
            out = out / cnt.view(-1, *([1] * (msg.dim() - 1)))
        elif self.aggr not in ("add", "sum"):
            raise NotImplementedError(self.aggr)
        return out

    def message(self, x_j):
        return x_j


class _RGCNConv(_MessagePassing):
    """torch_geometric.nn.RGCNConv stand-in (PyG 2.3.1, third-party), for the reference's comparison model `Net`
    (model.py:137-138): per-relation weights [R, in, out], root, bias; reset_parameters = glorot(weight), glorot(root),
    zeros(bias); forward without decomposition and without pyg_lib (not in the reference's requirements.txt) is the
    per-relation loop -- the very loop the reference's CustomRGCNConv was cut down from (mp_rgcn_layer.py:249-258):
        for i in range(R): h = propagate(edge_index[:, edge_type == i], x=x); out = out + h @ weight[i]
        out = out + x @ root + bias."""

    def __init__(self, in_channels, out_channels, num_relations, num_bases=None, num_blocks=None, aggr="mean",
                 root_weight=True, bias=True, **kwargs):
        kwargs.setdefault("aggr", aggr)
        super().__init__(node_dim=0, **kwargs)
        assert num_bases is None and num_blocks is None and root_weight
        self.in_channels, self.out_channels, self.num_relations = in_channels, out_channels, num_relations
        self.weight = torch.nn.Parameter(torch.empty(num_relations, in_channels, out_channels))
        self.root = torch.nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = torch.nn.Parameter(torch.empty(out_channels)) if bias else None
        _glorot(self.weight)
        _glorot(self.root)
        _zeros(self.bias)

    def forward(self, x, edge_index, edge_type=None):
        out = torch.zeros(x.size(0), self.out_channels, device=x.device)
        for i in range(self.num_relations):
            tmp = edge_index[:, edge_type == i]
            h = self.propagate(tmp, x=x, size=(x.size(0), x.size(0)))
            out = out + (h @ self.weight[i])
        out = out + x @ self.root
        if self.bias is not None:
            out = out + self.bias
        return out


def _glorot(t):
    if isinstance(t, torch.Tensor):
        a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        t.data.uniform_(-a, a)


def _zeros(t):
    if isinstance(t, torch.Tensor):
        t.data.fill_(0)


class _Comm:
    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0

    def bcast(self, obj, root=0):
        return obj

    def gather(self, obj, root=0):
        return [obj]


def _mod(name, **names):
    m = types.ModuleType(name)
    m.__dict__.update(names)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Register the stand-ins and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True

    class _SparseTensor:  # only used in isinstance() checks (mp_rgcn_layer.py:193)
        pass

    class _Dummy(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    tg = _mod("torch_geometric")
    tg.data = _mod("torch_geometric.data", Data=_Data)
    tg.nn = _mod("torch_geometric.nn", RGCNConv=_RGCNConv)
    tg.nn.conv = _mod("torch_geometric.nn.conv", MessagePassing=_MessagePassing)
    tg.nn.inits = _mod("torch_geometric.nn.inits", glorot=_glorot, zeros=_zeros)
    tg.typing = _mod("torch_geometric.typing", Adj=object, OptTensor=object)
    tg.loader = _mod("torch_geometric.loader", DataLoader=object)
    _mod("torch_scatter", scatter=None)
    _mod("torch_sparse", SparseTensor=_SparseTensor, masked_select_nnz=None, matmul=None)
    _mod("seaborn")
    mlx = _mod("mlxtend")
    mlx.plotting = _mod("mlxtend.plotting", plot_confusion_matrix=None)
    mpl = _mod("matplotlib")
    mpl.pyplot = _mod("matplotlib.pyplot")
    imb = _mod("imblearn")
    imb.under_sampling = _mod("imblearn.under_sampling", RandomUnderSampler=None)
    mpi = _mod("mpi4py")
    mpi.MPI = _mod("mpi4py.MPI", COMM_WORLD=_Comm())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def import_reference():
    """Return (main, model, mp_rgcn_layer) modules of the unmodified reference."""
    install()
    import main as ref_main  # noqa: E402  (reference main.py)
    import model as ref_model  # noqa: E402
    import mp_rgcn_layer as ref_layer  # noqa: E402

    ref_main.COMPLEX = "fb15k-237"  # main.py:1484 sets it only under __main__; read at :519
    return ref_main, ref_model, ref_layer
