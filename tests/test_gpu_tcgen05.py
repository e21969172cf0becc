"""tcgen05 (3xTF32) projection path against the exact-fp32 SIMT path and the oracle.
Bar: normalised max error <= 1e-5 (the fp32 parity bar of the north star).

The backward is compared on IDENTICAL saved tensors (x, h, y of one forward): ReLU is
discontinuous, so two fp32-accurate forwards may legitimately disagree on the sign of a
pre-activation that is ~1e-7 from zero, which would flip a whole gradient row."""
import pytest
import torch

from conftest import rel_err
from oracle import mpgnn_oracle as orc

import mpgnn_b200
from mpgnn_b200 import _lib

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]
DEV = "cuda"
TOL = 1e-5


def _graph(n, e, r, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (2, e), generator=g), torch.randint(0, r, (e,), generator=g)


def _fwd(graph, rel, x, w, root, b, flags, mask_bits, actmask=None):
    lib = _lib.load()
    n, f_in = x.shape
    f_out = w.size(1)
    h = torch.empty(n, f_in, device=DEV)
    y = torch.empty(n, f_out, device=DEV)
    ws = torch.empty(lib.mpgnn_hop_workspace_bytes(n, f_in, f_out), dtype=torch.uint8, device=DEV)
    _lib.check(lib.mpgnn_hop_fwd(graph.handle, rel, _lib.ptr(x), f_in, _lib.ptr(w), _lib.ptr(root), _lib.ptr(b), f_out,
                                 flags, 0.6, 0, 0, _lib.ptr(mask_bits), _lib.ptr(h), _lib.ptr(y), _lib.ptr(actmask),
                                 _lib.ptr(ws), ws.numel(), _lib.current_stream()))
    torch.cuda.synchronize()
    return h, y


def _unpack_actmask(words, f_out):
    """[n, f_out/32] int32 words -> [n, f_out] bool (bit j of word c = column 32c+j)."""
    sh = torch.arange(32, device=words.device, dtype=torch.int64)
    return ((words.to(torch.int64).unsqueeze(-1) >> sh) & 1).bool().reshape(words.size(0), f_out)


def _bwd(graph, rel, x, h, y, gy, w, root, flags, actmask=None, ws_fill=None):
    lib = _lib.load()
    n, f_in = x.shape
    f_out = w.size(1)
    gx = torch.empty(n, f_in, device=DEV)
    gw, gr, gb = torch.empty_like(w), torch.empty_like(root), torch.empty(f_out, device=DEV)
    ws = torch.empty(lib.mpgnn_hop_workspace_bytes(n, f_in, f_out), dtype=torch.uint8, device=DEV)
    if ws_fill is not None:
        ws.fill_(ws_fill)
    _lib.check(lib.mpgnn_hop_bwd(graph.handle, rel, _lib.ptr(x), _lib.ptr(h), _lib.ptr(y), _lib.ptr(actmask), _lib.ptr(gy), f_in,
                                 _lib.ptr(w), _lib.ptr(root), f_out, flags | _lib.F_NEED_GX, 0.6, _lib.ptr(gx),
                                 _lib.ptr(gw), _lib.ptr(gr), _lib.ptr(gb), _lib.ptr(ws), ws.numel(),
                                 _lib.current_stream()))
    torch.cuda.synchronize()
    return gx, gw, gr, gb


@pytest.mark.parametrize("n,f_in,f_out", [(128, 64, 64), (1000, 64, 64), (5000, 128, 128), (40000, 128, 128),
                                          (19001, 64, 128), (3000, 32, 64), (2500, 128, 64), (50001, 96, 192)])
def test_hop_tf32x3_matches_fp32_paths(n, f_in, f_out):
    from mpgnn_b200.mp_rgcn_layer import pack_mask_bits
    e, r = 6 * n, 3
    ei, et = _graph(n, e, r, seed=n)
    gen = torch.Generator().manual_seed(n + 1)
    x = torch.randn(n, f_in, generator=gen)
    gy = torch.randn(n, f_out, generator=gen).to(DEV)
    mask = (torch.rand(n, f_out, generator=gen) > 0.6).float()
    torch.manual_seed(n)
    p = orc.conv_init(f_in, f_out)
    w, root = p["weight"], p["root"]
    b = torch.randn(f_out, generator=gen) * 0.1
    graph = mpgnn_b200.RelationGraph(ei, et, n, r, device=DEV)
    xd, wd, rd, bd = x.to(DEV), w.to(DEV), root.to(DEV), b.to(DEV)
    bits = pack_mask_bits(mask.to(DEV))
    base = _lib.F_RELU | _lib.F_DROPOUT_MASK
    h32, y32 = _fwd(graph, 1, xd, wd, rd, bd, base, bits)
    htc, ytc = _fwd(graph, 1, xd, wd, rd, bd, base | _lib.F_TF32X3, bits)
    assert torch.equal(h32, htc)
    assert rel_err(ytc, y32) < TOL
    z, h, cnt = orc.conv_forward(x, ei, et, 1, w, root, b)
    assert rel_err(ytc, torch.relu(z) * mask * 2.5) < TOL
    g32 = _bwd(graph, 1, xd, h32, y32, gy, wd, rd, base)
    gtc = _bwd(graph, 1, xd, h32, y32, gy, wd, rd, base | _lib.F_TF32X3)
    for a, c in zip(gtc, g32):
        assert rel_err(a, c) < TOL
    # oracle backward on the same activation pattern (y32 > 0)
    gz = gy.cpu() * (y32.cpu() > 0) * 2.5
    gx_ref, gw_ref, gr_ref, gb_ref = orc.conv_backward(x, ei, et, 1, w, root, h, cnt, gz)
    for a, c in zip(gtc, (gx_ref, gw_ref, gr_ref, gb_ref)):
        assert rel_err(a, c) < TOL


@pytest.mark.parametrize("n,f_in,f_out,tc", [(5000, 128, 128, True), (19001, 128, 64, True), (3000, 64, 64, True),
                                             (2500, 96, 96, False), (777, 128, 128, False)])
def test_activation_bitmask_replaces_y_in_the_backward(n, f_in, f_out, tc):
    """hop_fwd's bitmask == [y > 0]; hop_bwd fed the bitmask (no y, g_z never materialised on the
    tensor-core path) returns bit-for-bit what the y-based backward returns."""
    ei, et = _graph(n, 5 * n, 2, seed=n + 7)
    gen = torch.Generator().manual_seed(n)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 2, device=DEV)
    x = torch.randn(n, f_in, generator=gen).to(DEV)
    gy = torch.randn(n, f_out, generator=gen).to(DEV)
    w = (torch.randn(f_in, f_out, generator=gen) * 0.1).to(DEV)
    root = (torch.randn(f_in, f_out, generator=gen) * 0.1).to(DEV)
    b = (torch.randn(f_out, generator=gen) * 0.1).to(DEV)
    flags = _lib.F_RELU | _lib.F_DROPOUT_SEED | (_lib.F_TF32X3 if tc else 0)
    am = torch.full((n, f_out // 32), -1, dtype=torch.int32, device=DEV)
    h, y = _fwd(graph, 1, x, w, root, b, flags, None, actmask=am)
    assert torch.equal(_unpack_actmask(am, f_out), y > 0)
    assert 0.1 < float((y > 0).float().mean()) < 0.3          # relu (~half) x keep 0.4
    ref = _bwd(graph, 1, x, h, y, gy, w, root, flags)
    got = _bwd(graph, 1, x, h, None, gy, w, root, flags, actmask=am)
    for a, c in zip(got, ref):
        assert torch.equal(a, c)


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("n,e,f_in,f_out", [(40000, 12000, 128, 128), (40000, 12000, 64, 128), (9000, 200000, 128, 64),
                                            (12345, 4000, 96, 64)])
def test_backward_never_reads_what_it_does_not_write(n, e, f_in, f_out, tc):
    """The input-gradient projection writes its two halves to two tensors and the transposed aggregation accumulates
    into g_x in place, skipping the rows without incoming edges: with the workspace poisoned by NaN patterns the
    result must be bit for bit the one obtained on a zeroed workspace (sparse relation: most rows have no edge;
    dense relation: almost every row has), i.e. nothing is read before the call itself wrote it."""
    ei, et = _graph(n, e, 1, seed=n + e)
    gen = torch.Generator().manual_seed(n)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 1, device=DEV)
    x = torch.randn(n, f_in, generator=gen).to(DEV)
    gy = torch.randn(n, f_out, generator=gen).to(DEV)
    w = (torch.randn(f_in, f_out, generator=gen) * 0.1).to(DEV)
    root = (torch.randn(f_in, f_out, generator=gen) * 0.1).to(DEV)
    b = (torch.randn(f_out, generator=gen) * 0.1).to(DEV)
    flags = _lib.F_RELU | _lib.F_DROPOUT_SEED | (_lib.F_TF32X3 if tc else 0)
    am = torch.empty((n, f_out // 32), dtype=torch.int32, device=DEV)
    h, y = _fwd(graph, 0, x, w, root, b, flags, None, actmask=am)
    clean = _bwd(graph, 0, x, h, None, gy, w, root, flags, actmask=am, ws_fill=0)
    dirty = _bwd(graph, 0, x, h, None, gy, w, root, flags, actmask=am, ws_fill=0xFF)
    for a, c in zip(dirty, clean):
        assert bool(torch.isfinite(a).all())
        assert torch.equal(a, c)
    # and the input gradient is the oracle's
    gz = gy.cpu() * (y.cpu() > 0) * 2.5
    _, hh, cnt = orc.conv_forward(x.cpu(), ei, et, 0, w.cpu(), root.cpu(), b.cpu())
    refs = orc.conv_backward(x.cpu(), ei, et, 0, w.cpu(), root.cpu(), hh, cnt, gz)
    for a, c in zip(dirty, refs):        # g_x, g_W (its producers skip the all-zero rows of h), g_root, g_bias
        assert rel_err(a, c) < TOL


def test_hop_tf32x3_seeded_dropout_same_stream_as_fp32():
    n, f = 3000, 128
    ei, et = _graph(n, 20000, 2, seed=1)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 2, device=DEV)
    conv = mpgnn_b200.CustomRGCNConv(f, f, 1, flow="target_to_source", device=DEV)
    x = torch.randn(n, f, device=DEV)
    with torch.no_grad():
        a = conv.hop(0, x, graph, relu=True, dropout_p=0.6, seed=77, offset=3, precision="fp32")
        b = conv.hop(0, x, graph, relu=True, dropout_p=0.6, seed=77, offset=3, precision="tf32x3")
        z = conv.hop(0, x, graph, precision="fp32")
    # same keep decisions from the counter RNG (a pre-activation within rounding of zero may take either side of the relu)
    sure = z.abs() > 1e-5
    assert torch.equal((a != 0) & sure, (b != 0) & sure)
    assert 0.35 < float((a != 0).float().mean()) / float((z > 0).float().mean()) < 0.45      # p = 0.6 dropped
    assert rel_err(b, a) < TOL


def test_tf32x3_accuracy_is_fp32_class():
    """The split must beat plain TF32 by orders of magnitude: compare with a float64 product."""
    n, f = 8192, 128
    ei, et = _graph(n, 4 * n, 1, seed=2)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 1, device=DEV)
    conv = mpgnn_b200.CustomRGCNConv(f, f, 1, flow="target_to_source", device=DEV)
    x = torch.randn(n, f, device=DEV) * 3.0
    with torch.no_grad():
        y = conv.hop(0, x, graph, precision="tf32x3")
        h = torch.zeros(n, f, dtype=torch.float64, device=DEV)
        cnt = torch.zeros(n, dtype=torch.float64, device=DEV)
        rows, cols = ei[0].to(DEV), ei[1].to(DEV)
        h.index_add_(0, rows, x.double()[cols])
        cnt.index_add_(0, rows, torch.ones(rows.numel(), dtype=torch.float64, device=DEV))
        ref = (h / cnt.clamp(min=1).unsqueeze(1)) @ conv.weight.double() + x.double() @ conv.root.double() \
            + conv.bias.double()
    assert rel_err(y, ref) < 2e-6


@pytest.mark.parametrize("n,f_in,f_out", [(400000, 64, 64), (200000, 32, 192), (150000, 96, 192), (100000, 128, 128)])
def test_projection_pipeline_is_race_free_exact_operand_check(n, f_in, f_out):
    """Stress for the TMA -> converter -> TMEM -> MMA handoffs.  With integer features, W = 0 and root a
    0/1 selection matrix, y must equal x (column c of y = column c % f_in of x) EXACTLY: the tf32 hi part
    carries the integer and lo is zero, so any row that picks up a piece of another chunk (a stage reused
    before its readers were done, seen once as a missing proxy fence) shows up bit-for-bit.  Repeated, since
    such races hit a few rows per million."""
    ei, et = _graph(n, 3 * n, 2, seed=n)
    gen = torch.Generator().manual_seed(n)
    x = torch.randint(-8, 9, (n, f_in), generator=gen).float().to(DEV)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 2, device=DEV)
    sel = torch.zeros(f_in, f_out, device=DEV)
    sel[torch.arange(f_out, device=DEV) % f_in, torch.arange(f_out, device=DEV)] = 1.0
    zero = torch.zeros(f_in, f_out, device=DEV)
    b = torch.zeros(f_out, device=DEV)
    want_x = x[:, torch.arange(f_out, device=DEV) % f_in]
    for trial in range(6):
        _, y = _fwd(graph, 1, x, zero, sel, b, _lib.F_TF32X3, None)
        assert torch.equal(y, want_x), "x operand corrupted in trial %d: %d rows" % (trial, int((y != want_x).any(1).sum()))
    h, _ = _fwd(graph, 1, x, zero, sel, b, 0, None)
    want_h = h[:, torch.arange(f_out, device=DEV) % f_in]
    for trial in range(6):
        _, y = _fwd(graph, 1, x, sel, zero, b, _lib.F_TF32X3, None)
        assert float((y - want_h).abs().max()) <= 1e-5 * 8, "h operand corrupted in trial %d" % trial


def _classes_of(fn):
    """Kernel classes (ScopedTimer names) the library recorded while `fn` ran."""
    lib = _lib.load()
    lib.mpgnn_timing_reset()
    lib.mpgnn_timing_enable(1)
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        lib.mpgnn_timing_enable(0)
    return set(_lib.timing_collect())


@pytest.mark.parametrize("n,f_in,f_out,eligible", [(40000, 128, 128, True), (5000, 64, 64, True), (3000, 32, 64, "proj"),
                                                   (4000, 2, 64, False), (3000, 100, 64, False)])
def test_default_path_runs_the_tcgen05_kernels(n, f_in, f_out, eligible):
    """The reference-facing call conv(layer_num, relation, x, edge_index, edge_type) (mp_rgcn_layer.py:158-159) takes
    the tensor-core kernels by default wherever the shape allows, and the SIMT kernels only as the odd-shape fallback;
    which one ran is read from the library's own per-kernel-class timers, not inferred from the numbers."""
    ei, et = _graph(n, 5 * n, 3, seed=7)
    x = torch.randn(n, f_in, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(3)
    conv = mpgnn_b200.CustomRGCNConv(f_in, f_out, 1, flow="target_to_source", device=DEV)
    assert conv.precision == "tf32x3"
    eid, etd = ei.to(DEV), et.to(DEV)
    xi = x.to(DEV).requires_grad_(True)
    out = {}

    def run():
        y = conv(0, 1, xi, eid, etd)
        y.backward(torch.ones_like(y))
        out["y"] = y

    classes = _classes_of(run)
    tc = {"proj_fwd_tcgen05", "wgrad_tn_tcgen05", "dgrad_nt_tcgen05"}
    simt = {"proj_fwd_simt", "wgrad_tn_simt", "dgrad_nt_simt"}
    if eligible is True:
        assert tc <= classes and not (simt & classes), sorted(classes)
    elif eligible == "proj":        # projection shapes eligible, weight-gradient shape (32-wide operand) not
        assert {"proj_fwd_tcgen05", "dgrad_nt_tcgen05", "wgrad_tn_simt"} <= classes, sorted(classes)
    else:
        assert simt <= classes and not (tc & classes), sorted(classes)
    # and the default path meets the fp32 bar against the oracle
    w, root, b = conv.weight.detach().cpu(), conv.root.detach().cpu(), conv.bias.detach().cpu()
    z, _, _ = orc.conv_forward(x, ei, et, 1, w, root, b)
    assert rel_err(out["y"], z) <= TOL
    # precision="fp32" keeps every kernel on the SIMT path
    classes32 = _classes_of(lambda: conv.hop(1, xi.detach(), mpgnn_b200.graph_for(eid, etd, n, torch.device(DEV)),
                                             precision="fp32"))
    assert "proj_fwd_simt" in classes32 and not (tc & classes32), sorted(classes32)


def test_model_default_path_runs_the_tcgen05_kernels(fx3):
    """MPNetm at the reference's defaults (hidden 64): hidden layers and their gradients on tcgen05, the 2-wide input
    layer on the thin SIMT kernels."""
    torch.manual_seed(30)
    model = mpgnn_b200.MPNetm(2, 64, fx3["num_relations"], 64, 2, 1, [[1, 0]], device=DEV)
    assert model.precision == "tf32x3"
    x, ei, et = fx3["x"].to(DEV), fx3["edge_index"].to(DEV), fx3["edge_type"].to(DEV)

    def run():
        model.train()
        model(x, ei, et).sum().backward()

    classes = _classes_of(run)
    assert {"proj_fwd_tcgen05", "wgrad_tn_tcgen05", "dgrad_nt_tcgen05"} <= classes, sorted(classes)


def test_bf16_flag_and_dropout_without_relu_are_rejected():
    """MPGNN_F_BF16 is reserved (no kernel behind it) and dropout without relu cannot be differentiated from [y > 0]:
    both are refused with MPGNN_ENOTSUP instead of silently running something else."""
    n, f = 2048, 64
    ei, et = _graph(n, 4 * n, 2, seed=1)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 2, device=DEV)
    x = torch.randn(n, f, device=DEV)
    w, root, b = torch.randn(f, f, device=DEV), torch.randn(f, f, device=DEV), torch.zeros(f, device=DEV)
    with pytest.raises(NotImplementedError):
        _fwd(graph, 0, x, w, root, b, _lib.F_RELU | _lib.F_BF16, None)
    with pytest.raises(NotImplementedError):
        _fwd(graph, 0, x, w, root, b, _lib.F_DROPOUT_SEED | _lib.F_TF32X3, None)
    conv = mpgnn_b200.CustomRGCNConv(f, f, 1, flow="target_to_source", device=DEV)
    with pytest.raises(NotImplementedError):
        conv.hop(0, x, graph, relu=False, dropout_p=0.5)
    with pytest.raises(NotImplementedError):
        conv.hop(0, x, graph, relu=True, precision="bf16")


# ---- compact hop (h kept as one row per non-empty bucket) ----------------------------------------------------------
def _sparse_graph(n, r, seed, hub=True):
    """~0.3 edges per (relation, node) like C4, plus (hub) one target with 700 and one source with 900 edges of relation 1."""
    g = torch.Generator().manual_seed(seed)
    e = int(0.3 * n * r)
    ei = torch.randint(0, n, (2, e), generator=g)
    et = torch.randint(0, r, (e,), generator=g)
    if hub:
        a = torch.randint(0, n, (700,), generator=g)
        b = torch.randint(0, n, (900,), generator=g)
        ei = torch.cat([ei, torch.stack([torch.full((700,), 17), a]), torch.stack([b, torch.full((900,), 23)])], 1)
        et = torch.cat([et, torch.ones(1600, dtype=torch.long)])
    return ei, et


def _relation_rows(graph, rel):
    import ctypes
    lib = _lib.load()
    p, cnt = ctypes.c_void_p(), ctypes.c_int64()
    _lib.check(lib.mpgnn_graph_relation_rows(graph.handle, rel, ctypes.byref(p), ctypes.byref(cnt)))
    from mpgnn_b200.graph import _device_view
    return _device_view(p.value, cnt.value, graph.device).clone().long()


@pytest.mark.parametrize("n,f_in,f_out,rel", [(40000, 128, 128, 1), (40000, 128, 128, 0), (25000, 64, 64, 1), (30001, 64, 128, 2),
                                              (6000, 128, 128, 3), (4737, 128, 128, 1), (4767, 128, 128, 0)])
def test_compact_hop_matches_dense_hop_and_oracle(n, f_in, f_out, rel):
    from mpgnn_b200.mp_rgcn_layer import pack_mask_bits
    lib = _lib.load()
    ei, et = _sparse_graph(n, 4, seed=n + rel)
    if rel == 3:                                  # a relation without any edge
        et = et.clamp(max=2)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 4, device=DEV)
    gen = torch.Generator().manual_seed(n)
    x = torch.randn(n, f_in, generator=gen)
    gy = torch.randn(n, f_out, generator=gen).to(DEV)
    mask = (torch.rand(n, f_out, generator=gen) > 0.6).float()
    torch.manual_seed(n)
    p = orc.conv_init(f_in, f_out)
    w, root = p["weight"], p["root"]
    b = torch.randn(f_out, generator=gen) * 0.1
    xd, wd, rd, bd = x.to(DEV), w.to(DEV), root.to(DEV), b.to(DEV)
    bits = pack_mask_bits(mask.to(DEV))
    base = _lib.F_RELU | _lib.F_DROPOUT_MASK | _lib.F_TF32X3
    rows = _relation_rows(graph, rel)
    nnz = rows.numel()
    assert nnz == torch.unique(ei[0][et == rel]).numel() and torch.equal(rows.cpu(), torch.unique(ei[0][et == rel]))
    assert lib.mpgnn_hop_h_rows(graph.handle, rel, f_in, f_out, base | _lib.F_COMPACT_H) == nnz
    assert lib.mpgnn_hop_h_rows(graph.handle, rel, f_in, f_out, base | _lib.F_DENSE_H) == n
    assert lib.mpgnn_hop_h_rows(graph.handle, rel, f_in, f_out, base) == n        # small graph: the dense form by default
    am_d = torch.empty(n, f_out // 32, dtype=torch.int32, device=DEV)
    am_c = torch.empty_like(am_d)
    h_d, y_d = _fwd(graph, rel, xd, wd, rd, bd, base | _lib.F_DENSE_H, bits, am_d)
    h_c, y_c = _fwd(graph, rel, xd, wd, rd, bd, base | _lib.F_COMPACT_H, bits, am_c)
    assert torch.equal(h_c[:nnz], h_d[rows])                          # same per-bucket sums in the same order, bit for bit
    z, _, _ = orc.conv_forward(x, ei, et, rel, w, root, b)
    y_ref = torch.relu(z) * mask * 2.5
    assert rel_err(y_c, y_ref) <= TOL and rel_err(y_c, y_d) <= TOL
    sure = (z.abs() > 1e-4).to(DEV)
    assert torch.equal(_unpack_actmask(am_c, f_out) & sure, _unpack_actmask(am_d, f_out) & sure)
    # backward on identical saved tensors (the dense forward's bitmask): compact vs dense
    out_d = _bwd(graph, rel, xd, h_d, None, gy, wd, rd, base | _lib.F_DENSE_H, am_d)
    hc_buf = torch.full_like(h_d, float("nan"))                       # only the first nnz rows may be read
    hc_buf[:nnz] = h_c[:nnz]
    out_c = _bwd(graph, rel, xd, hc_buf, None, gy, wd, rd, base | _lib.F_COMPACT_H, am_d, ws_fill=0xFF)
    for name, a, c in zip(("gx", "gw", "groot", "gbias"), out_d, out_c):
        assert torch.isfinite(c).all(), name
        assert rel_err(c, a) <= TOL, (name, rel_err(c, a))
    # and against the oracle's autograd-equivalent backward
    gz = gy.cpu() * _unpack_actmask(am_d, f_out).cpu() * 2.5
    _, h_o, cnt = orc.conv_forward(x, ei, et, rel, w, root, b)
    ref = orc.conv_backward(x, ei, et, rel, w, root, h_o, cnt, gz)
    for name, a, c in zip(("gx", "gw", "groot", "gbias"), ref, out_c):
        assert rel_err(c, a) <= TOL, (name, rel_err(c, a))


def test_compact_hop_kernel_classes_and_autograd_path():
    """The compact form through the Python autograd function (flags carried from forward to backward) and the kernel
    classes it launches; shapes whose weight gradient has no single-operand tensor-core form stay dense."""
    n, f = 20000, 128
    ei, et = _sparse_graph(n, 4, seed=5, hub=False)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 4, device=DEV)
    torch.manual_seed(1)
    conv = mpgnn_b200.CustomRGCNConv(f, f, 1, flow="target_to_source", device=DEV)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(2))
    xi = x.to(DEV).requires_grad_(True)
    gy = torch.randn(n, f, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    out = {}

    def run():
        y = conv.hop(2, xi, graph, relu=True, h_layout="compact")
        y.backward(gy)
        out["y"] = y.detach()

    classes = _classes_of(run)
    want = {"spmm_mean_fwd", "proj_fwd_compact_tcgen05", "proj_fwd_tcgen05", "gather_gz_compact", "wgrad_tn_tcgen05",
            "wgrad_tn_compact_tcgen05", "dgrad_nt_tcgen05", "dgrad_nt_compact_tcgen05", "spmm_transpose_bwd"}
    assert want <= classes and not any(c.endswith("_simt") for c in classes), sorted(classes)
    grads_c = [xi.grad.clone(), conv.weight.grad.clone(), conv.root.grad.clone(), conv.bias.grad.clone()]
    xi.grad = None
    conv.zero_grad()
    y = conv.hop(2, xi, graph, relu=True, h_layout="dense")
    y.backward(gy)
    assert rel_err(out["y"], y) <= TOL
    for a, c in zip([xi.grad, conv.weight.grad, conv.root.grad, conv.bias.grad], grads_c):
        assert rel_err(c, a) <= 5 * TOL                      # two forwards: a relu within rounding of zero may differ
    lib = _lib.load()
    assert lib.mpgnn_hop_h_rows(graph.handle, 2, 128, 64, _lib.F_TF32X3 | _lib.F_COMPACT_H) == n      # no such wgrad shape
    assert lib.mpgnn_hop_h_rows(graph.handle, 2, 96, 128, _lib.F_TF32X3 | _lib.F_COMPACT_H) == n
    assert lib.mpgnn_hop_h_rows(graph.handle, 2, 128, 128, _lib.F_COMPACT_H) == n                     # fp32 (SIMT) path


@pytest.mark.parametrize("m", [1, 31, 32, 33, 148 * 32 - 1, 148 * 32, 148 * 32 + 1, 100003])
def test_gemm_tn_row_counts_around_chunk_and_cta_boundaries(m):
    """out = A^T B and colsum(B) of the dense helper (fixed-order split over the rows) against a float64 product, at row
    counts on both sides of the 32-row and 148-CTA boundaries; twice into NaN-filled outputs and scratch (no stale
    partial may be read, run-to-run bits identical).  The tensor-core weight gradients see the same row counts through
    the compact-hop cases above (n = 4737, 4767, 6000)."""
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(m)
    a = torch.randn(m, 128, device=DEV, generator=g)
    b = torch.randn(m, 128, device=DEV, generator=g)
    wsb = lib.mpgnn_gemm_workspace_bytes(m, 128, 128)
    ref = a.double().t() @ b.double()
    ref_cs = b.double().sum(0)
    outs = []
    for rep in range(2):
        out = torch.full((128, 128), float("nan"), device=DEV)
        cs = torch.full((128,), float("nan"), device=DEV)
        ws = torch.full((wsb // 4,), float("nan"), device=DEV)
        _lib.check(lib.mpgnn_gemm_tn(_lib.ptr(a), 128, m, 128, _lib.ptr(b), 128, 128, _lib.ptr(out), 128, _lib.ptr(cs),
                                     _lib.ptr(ws), wsb, _lib.current_stream()))
        torch.cuda.synchronize()
        assert rel_err(out.double(), ref) <= TOL and rel_err(cs.double(), ref_cs) <= TOL
        outs.append((out, cs))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
