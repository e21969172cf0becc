"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs, against the committed reference goldens, and through size-independent properties
at larger sizes.  Bars: bit-exact for CSR / relation indexing; normalised max error
<= 1e-5 (north star: 1e-5 relative, fp32) for floating point."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import mpgnn_oracle as orc

import mpgnn_b200
from mpgnn_b200 import _lib

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
DEV = "cuda"


def _rand_graph(n, e, r, seed, dup=True):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    et = torch.randint(0, r, (e,), generator=g)
    if dup and e >= 8:  # force duplicate triplets and a self loop
        ei[:, 1] = ei[:, 0]
        et[1] = et[0]
        ei[1, 2] = ei[0, 2]
    return ei, et


def _check_graph(ei, et, n, r, from_host):
    ptr, col, perm = orc.relation_csr(ei.numpy(), et.numpy(), n, r)
    ptr_t, row_t, perm_t = orc.relation_csr(ei.numpy(), et.numpy(), n, r, transpose=True)
    g = mpgnn_b200.RelationGraph(ei if from_host else ei.to(DEV), et if from_host else et.to(DEV), n, r, device=DEV)
    assert np.array_equal(g.relation_counts, np.bincount(et.numpy(), minlength=r))
    for rel in range(r):
        for tr, (p_o, i_o, e_o) in ((False, (ptr, col, perm)), (True, (ptr_t, row_t, perm_t))):
            p, i, e = g.relation_view(rel, transpose=tr)
            lo, hi = p_o[rel * n], p_o[(rel + 1) * n]
            assert np.array_equal(p.cpu().numpy(), (p_o[rel * n:(rel + 1) * n + 1] - lo).astype(np.int32))
            assert np.array_equal(i.cpu().numpy(), i_o[lo:hi])
            assert np.array_equal(e.cpu().numpy(), e_o[lo:hi])  # stable: original edge order inside a bucket
    return g


def test_k1_csr_fixtures_bit_exact(fx3, fx4):
    for fx in (fx3, fx4):
        _check_graph(fx["edge_index"], fx["edge_type"], fx["x"].size(0), fx["num_relations"], from_host=True)


@pytest.mark.parametrize("n,e,r", [(1, 0, 1), (1, 5, 1), (7, 1, 3), (257, 4096, 5), (1000, 4097, 64),
                                   (300, 70000, 237), (50000, 200000, 20)])
def test_k1_csr_random_bit_exact(n, e, r):
    ei, et = _rand_graph(n, e, r, seed=n + e + r)
    _check_graph(ei, et, n, r, from_host=(e % 2 == 0))


def test_k1_rejects_out_of_range():
    ei, et = _rand_graph(10, 50, 3, 1)
    bad = et.clone()
    bad[7] = 3
    with pytest.raises(ValueError):
        mpgnn_b200.RelationGraph(ei, bad, 10, 3, device=DEV)
    bad_ei = ei.clone()
    bad_ei[1, 3] = 10
    with pytest.raises(ValueError):
        mpgnn_b200.RelationGraph(bad_ei, et, 10, 3, device=DEV)


def _spmm(g, rel, x, transpose=False, mean=True, init=None):
    lib = _lib.load()
    out = torch.empty(x.size(0), x.size(1), device=DEV)
    rc = lib.mpgnn_spmm(g.handle, rel, int(transpose), int(mean), _lib.ptr(x), x.stride(0), x.size(1),
                        _lib.ptr(init), 0 if init is None else init.stride(0), _lib.ptr(out), out.stride(0),
                        _lib.current_stream())
    _lib.check(rc)
    return out


@pytest.mark.parametrize("feat", [1, 2, 3, 6, 8, 20, 64, 100, 128, 256, 260])
def test_k2_spmm_mean_matches_oracle(feat):
    n, e, r = 3000, 20000, 4
    ei, et = _rand_graph(n, e, r, seed=feat)
    ei[0, :200] = 5  # one high-degree row (>32 edges, below the hub threshold): exercises the batched index path
    et[:200] = 1
    g = mpgnn_b200.RelationGraph(ei, et, n, r, device=DEV)
    x = torch.randn(n, feat, generator=torch.Generator().manual_seed(0))
    for rel in range(r):
        ref, _ = orc.propagate_mean(x, orc.masked_edge_index(ei, et == rel))
        out = _spmm(g, rel, x.to(DEV))
        assert rel_err(out, ref) < 1e-6
    # edge-order summation: equal to a sequential fp32 loop, bit for bit
    tmp = orc.masked_edge_index(ei, et == 1).numpy()
    acc = np.zeros((n, feat), np.float32)
    xn = x.numpy()
    for a, b in zip(tmp[0], tmp[1]):
        acc[a] += xn[b]
    deg = np.maximum(np.bincount(tmp[0], minlength=n), 1).astype(np.float32)
    assert np.array_equal(_spmm(g, 1, x.to(DEV)).cpu().numpy(), acc / deg[:, None])


@pytest.mark.parametrize("feat", [64, 128, 260])
def test_k2_spmm_hub_buckets_chunked_path(feat):
    """Power-law shape: buckets with more than 256 edges (hub targets in the CSR view, hub sources in the CSC view)
    are summed in chunks by separate warps and combined in chunk order.  Checked against a float64 gather (the
    chunked tree is not the sequential fp32 order any more), for mean / sum / init, run to run identical."""
    n, e, r = 6000, 60000, 3
    gen = torch.Generator().manual_seed(feat)
    ei = torch.randint(0, n, (2, e), generator=gen)
    et = torch.randint(0, r, (e,), generator=gen)
    ei[0, :9000] = 17                      # hub target of relation 1: 9000 in-edges -> 36 chunks
    ei[0, 9000:9300] = 4000                # just above the threshold: 2 chunks
    et[:9300] = 1
    ei[1, 20000:31000] = 123               # hub source of relation 2 (transposed view): 11000 edges
    et[20000:31000] = 2
    g = mpgnn_b200.RelationGraph(ei, et, n, r, device=DEV)
    x = torch.randn(n, feat, generator=gen)
    init = torch.randn(n, feat, generator=gen)
    xd, initd = x.to(DEV), init.to(DEV)
    for rel in range(r):
        m = et == rel
        rows, cols = ei[0][m], ei[1][m]
        for transpose in (False, True):
            tgt, src = (cols, rows) if transpose else (rows, cols)
            ref = torch.zeros(n, feat, dtype=torch.float64).index_add_(0, tgt, x.double()[src])
            deg = torch.bincount(tgt, minlength=n).clamp(min=1).double().unsqueeze(1)
            out_sum = _spmm(g, rel, xd, transpose=transpose, mean=False)
            out_mean = _spmm(g, rel, xd, transpose=transpose, mean=True)
            out_init = _spmm(g, rel, xd, transpose=transpose, mean=False, init=initd)
            assert rel_err(out_sum, ref.float()) < 2e-6
            assert rel_err(out_mean, (ref / deg).float()) < 2e-6
            assert rel_err(out_init, (ref + init.double()).float()) < 2e-6
            assert torch.equal(out_sum, _spmm(g, rel, xd, transpose=transpose, mean=False))     # deterministic


def _spmm_inplace(g, rel, x, buf, transpose=False, mean=False):
    lib = _lib.load()
    rc = lib.mpgnn_spmm(g.handle, rel, int(transpose), int(mean), _lib.ptr(x), x.stride(0), x.size(1),
                        _lib.ptr(buf), buf.stride(0), _lib.ptr(buf), buf.stride(0), _lib.current_stream())
    _lib.check(rc)
    return buf


@pytest.mark.parametrize("feat", [8, 32, 36, 64, 128, 256, 260])
@pytest.mark.parametrize("shape", ["sparse", "dense", "hub"])
def test_k2_spmm_in_place_equals_out_of_place(feat, shape):
    """d_init aliasing d_out (the backward's g_x += A^T t): rows without edges are neither read nor written and the
    others start from the stored value and add their bucket in edge order -- bit for bit the out-of-place result,
    for the shapes that take the skipping kernel (feat % 4 == 0, >= 32) and those that do not."""
    gen = torch.Generator().manual_seed(feat)
    if shape == "sparse":          # most buckets empty, like one relation of C4; n not a multiple of 32
        n, e, r = 50021, 30000, 3
    elif shape == "dense":         # degrees around 40: several index batches per row at every group width
        n, e, r = 1500, 120000, 2
    else:
        n, e, r = 6000, 60000, 3
    ei = torch.randint(0, n, (2, e), generator=gen)
    et = torch.randint(0, r, (e,), generator=gen)
    if shape == "hub":
        ei[1, :9000] = 17          # hub source in the transposed view (36 chunks) ...
        ei[0, 9000:9300] = 4000    # ... and a hub target in the forward view
        et[:9300] = 1
    g = mpgnn_b200.RelationGraph(ei, et, n, r, device=DEV)
    x = torch.randn(n, feat, generator=gen).to(DEV)
    init = torch.randn(n, feat, generator=gen).to(DEV)
    for rel in range(r):
        for transpose in (False, True):
            for mean in (False, True):
                ref = _spmm(g, rel, x, transpose=transpose, mean=mean, init=init)
                got = _spmm_inplace(g, rel, x, init.clone(), transpose=transpose, mean=mean)
                assert torch.equal(got, ref), (rel, transpose, mean)
    # a strided destination (columns of a wider tensor) stays inside its columns
    wide = torch.randn(n, 2 * feat + 4, generator=gen).to(DEV)
    if (feat % 4) == 0:
        keep = wide.clone()
        view = wide[:, feat:2 * feat]
        ref = _spmm(g, 1, x, transpose=True, mean=False, init=view)
        _spmm_inplace(g, 1, x, view, transpose=True)
        assert torch.equal(view, ref)
        assert torch.equal(wide[:, :feat], keep[:, :feat]) and torch.equal(wide[:, 2 * feat:], keep[:, 2 * feat:])


def test_k2_spmm_transpose_is_adjoint():
    n, e, r, f = 20000, 150000, 8, 64
    ei, et = _rand_graph(n, e, r, seed=3)
    g = mpgnn_b200.RelationGraph(ei.to(DEV), et.to(DEV), n, r)
    gen = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(n, f, device=DEV, generator=gen)
    y = torch.randn(n, f, device=DEV, generator=gen)
    for rel in (0, 5):
        ax = _spmm(g, rel, x, mean=False)
        a = (ax * y).double().sum()
        b = (x * _spmm(g, rel, y, transpose=True, mean=False)).double().sum()
        # <Ax, y> == <x, A^T y> up to fp32 rounding of the two gathers (Cauchy-Schwarz scale)
        assert abs(float(a - b)) < 1e-6 * float(ax.double().norm() * y.double().norm())
    # init accumulation
    o = _spmm(g, 2, x, transpose=True, mean=False, init=y)
    assert rel_err(o, y + _spmm(g, 2, x, transpose=True, mean=False)) < 1e-6


@pytest.mark.parametrize("tag", ["l0", "l1"])
def test_layer_forward_backward_matches_reference_golden(fx3, tag):
    g = load_golden("layer_len3")
    s = int(g["row_stride"])
    x = fx3["x"] if tag == "l0" else torch.from_numpy(g["x64"])
    g_out = torch.from_numpy(g["g64"]).to(DEV)
    conv = mpgnn_b200.CustomRGCNConv(x.size(1), 64, 1, flow="target_to_source", device=DEV)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(g[tag + "_weight"]))
        conv.root.copy_(torch.from_numpy(g[tag + "_root"]))
        conv.bias.copy_(torch.from_numpy(g[tag + "_bias"]))
    for r in range(fx3["num_relations"]):
        xi = x.clone().to(DEV).requires_grad_(True)
        conv.zero_grad()
        out = conv(0, r, xi, fx3["edge_index"], fx3["edge_type"])  # the reference call shape
        (out * g_out).sum().backward()
        pre = "%s_r%d_" % (tag, r)
        assert rel_err(out[::s], g[pre + "out"]) < FP32_TOL
        assert rel_err(xi.grad[::s], g[pre + "gx"]) < FP32_TOL
        assert rel_err(conv.weight.grad, g[pre + "gw"]) < FP32_TOL
        assert rel_err(conv.root.grad, g[pre + "groot"]) < FP32_TOL
        assert rel_err(conv.bias.grad, g[pre + "gbias"]) < FP32_TOL


@pytest.mark.parametrize("n,e,r,f_in,f_out", [(1, 0, 1, 4, 4), (33, 200, 2, 2, 64), (5000, 30000, 4, 64, 64),
                                              (4100, 60000, 6, 100, 64), (2049, 9000, 3, 128, 128),
                                              (777, 5000, 2, 7, 5),
                                              (14541, 272115, 237, 100, 64)])     # BASELINE configs[2] (FB15K-237) shape
def test_hop_fused_epilogue_matches_oracle(n, e, r, f_in, f_out):
    ei, et = _rand_graph(n, e, r, seed=n)
    gen = torch.Generator().manual_seed(n)
    x = torch.randn(n, f_in, generator=gen)
    gy = torch.randn(n, f_out, generator=gen)
    mask = (torch.rand(n, f_out, generator=gen) > 0.6).float()
    conv = mpgnn_b200.CustomRGCNConv(f_in, f_out, 1, flow="target_to_source", device=DEV)
    with torch.no_grad():
        conv.bias.copy_(torch.randn(f_out, generator=gen) * 0.1)
    w, root, b = conv.weight.detach().cpu(), conv.root.detach().cpu(), conv.bias.detach().cpu()
    graph = mpgnn_b200.RelationGraph(ei, et, n, r, device=DEV)
    rel = int(torch.bincount(et, minlength=r).argmax()) if r > 64 else r - 1     # a populated relation of the big one
    z, h, cnt = orc.conv_forward(x, ei, et, rel, w, root, b)
    y_ref = torch.relu(z) * mask * 2.5
    xi = x.to(DEV).requires_grad_(True)
    y = conv.hop(rel, xi, graph, relu=True, dropout_p=0.6, dropout_mask=mask)
    y.backward(gy.to(DEV))
    assert rel_err(y, y_ref) < FP32_TOL
    # the backward is compared on the device's own activation pattern: an element with |z| ~ 1e-8 may sit on either
    # side of the ReLU in two correct fp32 evaluations (the CPU matmul is not bit-reproducible run to run), and one
    # flipped gate moves g_x by far more than the tolerance
    act = y.detach().cpu() > 0
    assert bool(((act != (y_ref > 0)) <= (z.abs() < 1e-5)).all())
    gz = gy * act * 2.5
    gx_ref, gw_ref, gr_ref, gb_ref = orc.conv_backward(x, ei, et, rel, w, root, h, cnt, gz)
    assert rel_err(xi.grad, gx_ref) < FP32_TOL
    assert rel_err(conv.weight.grad, gw_ref) < FP32_TOL
    assert rel_err(conv.root.grad, gr_ref) < FP32_TOL
    assert rel_err(conv.bias.grad, gb_ref) < FP32_TOL


def test_hop_seeded_dropout_statistics_and_determinism():
    n, f = 4096, 64
    ei, et = _rand_graph(n, 20000, 2, seed=9)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 2, device=DEV)
    conv = mpgnn_b200.CustomRGCNConv(f, f, 1, flow="target_to_source", device=DEV)
    x = torch.randn(n, f, device=DEV)
    with torch.no_grad():
        base = conv.hop(0, x, graph, relu=True)
        a = conv.hop(0, x, graph, relu=True, dropout_p=0.6, seed=123, offset=5)
        b = conv.hop(0, x, graph, relu=True, dropout_p=0.6, seed=123, offset=5)
        c = conv.hop(0, x, graph, relu=True, dropout_p=0.6, seed=124, offset=5)
    assert torch.equal(a, b) and not torch.equal(a, c)
    pos = base > 0
    kept = (a != 0) & pos
    assert abs(float(kept.sum()) / float(pos.sum()) - 0.4) < 0.01
    assert rel_err(a[kept], base[kept] * 2.5) < 1e-6


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def _model_from(sd, metapaths, precision="tf32x3"):
    m = mpgnn_b200.MPNetm(2, 64, 4, 64, 2, len(metapaths), metapaths, device=DEV, precision=precision)
    m.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
    return m


@pytest.mark.parametrize("precision", ["tf32x3", "fp32"])
def test_model_eval_and_train_step_match_reference_golden(fx3, precision):
    """Both projection paths against the reference's recorded forward, loss, gradients and first Adam step.  Logits and
    loss meet the 1e-5 bar on both.  Gradients: the exact-fp32 path (same summation orders as torch) meets 5e-5
    (measured 1e-7).  On the 3xTF32 path everything behind a relu is held to 5e-3: of the 3 x 320,000 pre-activations
    (two conv layers and fc1) about one lies within rounding of zero, an fp32-accurate product may put it on the other
    side of the relu than torch did, and that unit's node then moves its share of every gradient upstream of it
    (~1/3600 of the loss x |W| -- measured 3e-4 .. 1.2e-3 of the largest entry); fc2's gradient does not pass through
    a gate and stays at 5e-5 (measured 2e-6).  test_gpu_tcgen05.py compares the two paths' backward at 1e-5 on
    identical saved activations, which is the statement about the kernels' accuracy."""
    g = load_golden("model_len3")
    data = mpgnn_b200.Data(**{k: fx3[k] for k in ("x", "edge_index", "edge_type", "train_idx", "train_y", "val_idx",
                                                   "val_y", "test_idx", "test_y")}, num_nodes=fx3["x"].size(0))
    model = _model_from(_sd(g, "sd0."), [[1, 0]], precision)
    model.eval()
    with torch.no_grad():
        logp = model(fx3["x"], fx3["edge_index"], fx3["edge_type"])
    assert rel_err(logp, g["eval_logp"]) < FP32_TOL
    masks = {(0, k): torch.from_numpy(np.unpackbits(g["step_mask_%d" % k], axis=1)[:, :64].astype(np.float32))
             for k in range(2)}
    model.inject_dropout_masks(masks)
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0005)
    loss, _ = mpgnn_b200.mpgnn_train(model, opt, data)
    assert abs(loss - float(g["step_loss"])) < FP32_TOL * abs(float(g["step_loss"]))
    errs = {k: rel_err(p.grad, g["step_grad." + k]) for k, p in model.named_parameters()}
    print(precision, {k: "%.1e" % v for k, v in errs.items()})
    for k, p in model.named_parameters():
        tol = 5e-3 if (precision == "tf32x3" and not k.startswith("fc2")) else 5 * FP32_TOL
        assert errs[k] < tol, (k, errs[k])
        assert rel_err(p.detach(), g["sd1." + k]) < (FP32_TOL if precision == "fp32" else 1e-4), k
    model.inject_dropout_masks(None)
    f1_tr, f1_va, _, loss_val = mpgnn_b200.mpgnn_validation(model, data, None)
    assert abs(f1_tr - g["step_val"][0]) < 2e-3 and abs(f1_va - g["step_val"][1]) < 2e-3
    assert abs(float(loss_val) - g["step_val"][2]) < 1e-4 * abs(g["step_val"][2])


def test_training_trace_matches_reference_golden(fx3):
    g = load_golden("model_len3")
    data = mpgnn_b200.Data(**{k: fx3[k] for k in ("x", "edge_index", "edge_type", "train_idx", "train_y", "val_idx",
                                                   "val_y", "test_idx", "test_y")}, num_nodes=fx3["x"].size(0))
    model = _model_from(_sd(g, "sd0."), [[1, 0]])
    model.dropout.p = 0.0
    model.dropout2.p = 0.0
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0005)
    ref = g["trace20_m10"]
    for ep in range(20):
        loss, cw = mpgnn_b200.mpgnn_train(model, opt, data)
        f1_tr, f1_va, _, loss_val = mpgnn_b200.mpgnn_validation(model, data, cw)
        assert abs(loss - ref[ep, 0]) < 1e-4 * abs(ref[ep, 0]), ep
        assert abs(float(loss_val) - ref[ep, 1]) < 1e-4 * abs(ref[ep, 1]), ep
        assert abs(f1_tr - ref[ep, 2]) < 2e-3 and abs(f1_va - ref[ep, 3]) < 2e-3, ep
    for k, p in model.named_parameters():
        assert rel_err(p.detach(), g["sd20." + k]) < 1e-3, k
    loss_t, f1_t = mpgnn_b200.mpgnn_test(model, data, None)
    assert abs(float(loss_t) - g["trace20_m10_test"][0]) < 1e-3 * abs(g["trace20_m10_test"][0])
    assert abs(f1_t - g["trace20_m10_test"][1]) < 2e-3


def test_macro_f1_kernel_matches_oracle():
    rng = np.random.RandomState(0)
    for c in (2, 3, 7):
        logp = torch.from_numpy(rng.randn(500, c).astype(np.float32)).to(DEV)
        idx = torch.from_numpy(rng.permutation(500)[:300].astype(np.int64)).to(DEV)
        y = torch.from_numpy(rng.randint(0, c, 300).astype(np.int64)).to(DEV)
        f1 = float(mpgnn_b200.device_macro_f1(logp, idx, y).item())
        ref = orc.macro_f1(logp[idx].argmax(1).cpu().numpy(), y.cpu().numpy())
        assert abs(f1 - ref) < 1e-12
    logp = torch.zeros(10, 3, device=DEV)
    logp[:, 1] = 1.0  # every prediction is class 1, truth is class 1: labels present = {1}
    idx = torch.arange(10, device=DEV)
    assert float(mpgnn_b200.device_macro_f1(logp, idx, torch.ones(10, dtype=torch.int64, device=DEV)).item()) == 1.0


def test_adam_and_nll_kernels_match_oracle():
    lib = _lib.load()
    gen = torch.Generator().manual_seed(0)
    p = torch.randn(1000, generator=gen)
    sd, state = {"p": p.clone()}, {}
    pd, m, v = p.clone().to(DEV), torch.zeros(1000, device=DEV), torch.zeros(1000, device=DEV)
    for step in range(1, 6):
        grad = torch.randn(1000, generator=gen)
        sd = orc.adam_step(sd, {"p": grad}, state)
        grad_d = grad.to(DEV)  # keep the device copy alive across the asynchronous call
        _lib.check(lib.mpgnn_adam_step(_lib.ptr(pd), _lib.ptr(grad_d), _lib.ptr(m), _lib.ptr(v), 1000, step,
                                       0.01, 0.9, 0.999, 1e-8, 0.0005, _lib.current_stream()))
        assert rel_err(pd, sd["p"]) < 1e-6
    n, c = 3000, 5
    logits = torch.randn(n, c, generator=gen)
    idx = torch.randperm(n, generator=gen)[:1700]
    y = torch.randint(0, c, (1700,), generator=gen)
    lg = logits.clone().requires_grad_(True)
    ref_logp = torch.log_softmax(lg, 1)
    ref_loss = orc.nll_loss_on_index(ref_logp, idx, y)
    ref_loss.backward()
    logp = torch.empty(n, c, device=DEV)
    loss = torch.empty(1, device=DEV)
    glog = torch.empty(n, c, device=DEV)
    ws = torch.empty(1 << 16, dtype=torch.uint8, device=DEV)
    logits_d, idx_d, y_d = logits.to(DEV), idx.to(DEV), y.to(DEV)
    _lib.check(lib.mpgnn_logsoftmax_nll(_lib.ptr(logits_d), n, c, _lib.ptr(idx_d), _lib.ptr(y_d),
                                        1700, _lib.ptr(logp), _lib.ptr(loss), _lib.ptr(glog), _lib.ptr(ws),
                                        ws.numel(), _lib.current_stream()))
    assert rel_err(logp, ref_logp.detach()) < 1e-6
    assert abs(float(loss) - float(ref_loss)) < 1e-6 * abs(float(ref_loss))
    assert rel_err(glog, lg.grad) < 1e-5


def test_hop_linearity_at_scale():
    """Size-independent property at a size the oracle would not finish quickly:
    hop(a*x1 + b*x2) = a*hop(x1) + b*hop(x2) (no bias, no activation)."""
    n, e, r, f = 400000, 4000000, 16, 128
    g = torch.Generator(device=DEV).manual_seed(5)
    ei = torch.randint(0, n, (2, e), device=DEV, generator=g)
    et = torch.randint(0, r, (e,), device=DEV, generator=g)
    graph = mpgnn_b200.RelationGraph(ei, et, n, r)
    assert int(graph.relation_counts.sum()) == e
    conv = mpgnn_b200.CustomRGCNConv(f, f, 1, flow="target_to_source", device=DEV)
    x1 = torch.randn(n, f, device=DEV, generator=g)
    x2 = torch.randn(n, f, device=DEV, generator=g)
    with torch.no_grad():
        y1, y2 = conv.hop(3, x1, graph), conv.hop(3, x2, graph)
        y12 = conv.hop(3, 0.5 * x1 - 2.0 * x2, graph)
    assert rel_err(y12, 0.5 * y1 - 2.0 * y2) < FP32_TOL
    # CSR sortedness / completeness: ptr is monotone and covers exactly E_r edges
    p, i, eid = graph.relation_view(3)
    assert bool((p[1:] >= p[:-1]).all()) and int(p[-1]) == graph.relation_edges(3)
    assert bool((et[eid.long()] == 3).all()) and bool((ei[1][eid.long()] == i.long()).all())
    rows = torch.repeat_interleave(torch.arange(n, device=DEV), (p[1:] - p[:-1]).long())
    assert bool((ei[0][eid.long()] == rows).all())
    same_row = rows[1:] == rows[:-1]
    assert bool((eid[1:][same_row] > eid[:-1][same_row]).all())  # stable inside each bucket


def _dot64(a, b, block=1 << 20):
    """<a, b> accumulated in float64 block by block (the operands are 5 GB each)."""
    tot = 0.0
    a2, b2 = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    for i in range(0, a2.size(0), block):
        tot += float((a2[i:i + block].double() * b2[i:i + block].double()).sum())
    return tot


@pytest.mark.timeout(600, method="thread")
def test_c4_full_size_properties():
    """BASELINE.json configs[3] at FULL size (10M nodes / 200M edges / 64 relations / hidden 128), through the
    properties the domain offers instead of the oracle:
      * K1: every edge lands in exactly one bucket (counts = bincount of the types; ptr monotone; eids of a sampled
        relation are a stable selection of its edges);
      * forward (K2 + K3, tensor-core path): linear in x for fixed weights;
      * backward (K4, dgrad, transposed K2): the adjoint identities of the bilinear map y = mean_r(x) W + x root + b:
        <y(x), g> = <x, g_x>  and  <y, g> = <W, g_W> + <root, g_root> + <b, g_b>."""
    free, _ = torch.cuda.mem_get_info()
    if free < 70 * 2 ** 30:
        pytest.skip("needs ~60 GB of free HBM")
    n, e, r, f = 10_000_000, 200_000_000, 64, 128
    g = torch.Generator(device=DEV).manual_seed(0)
    ei = torch.randint(0, n, (2, e), device=DEV, generator=g)
    et = torch.randint(0, r, (e,), device=DEV, generator=g)
    graph = mpgnn_b200.RelationGraph(ei, et, n, r)
    counts = torch.bincount(et, minlength=r).cpu().numpy()
    assert np.array_equal(np.asarray(graph.relation_counts), counts) and int(counts.sum()) == e
    rel = 37
    p, idx, eid = graph.relation_view(rel)
    assert bool((p[1:] >= p[:-1]).all()) and int(p[-1]) == int(counts[rel])
    assert bool((et[eid.long()] == rel).all()) and bool((ei[1][eid.long()] == idx.long()).all())
    rows = torch.repeat_interleave(torch.arange(n, device=DEV), (p[1:] - p[:-1]).long())
    assert bool((ei[0][eid.long()] == rows).all())
    same_row = rows[1:] == rows[:-1]
    assert bool((eid[1:][same_row] > eid[:-1][same_row]).all())            # stable inside each bucket
    del ei, et, rows, same_row
    conv = mpgnn_b200.CustomRGCNConv(f, f, 1, flow="target_to_source", device=DEV)
    with torch.no_grad():
        conv.bias.copy_(torch.randn(f, device=DEV, generator=g) * 0.1)
    x1 = torch.randn(n, f, device=DEV, generator=g)
    x2 = torch.randn(n, f, device=DEV, generator=g)
    with torch.no_grad():
        zero = conv.hop(rel, torch.zeros(1, f, device=DEV).expand(n, f), graph, precision="tf32x3")   # = bias rows
        y1 = conv.hop(rel, x1, graph, precision="tf32x3")
        y2 = conv.hop(rel, x2, graph, precision="tf32x3")
        x2.mul_(-2.0).add_(x1, alpha=0.5)                                  # x2 <- 0.5 x1 - 2 x2
        y12 = conv.hop(rel, x2, graph, precision="tf32x3")
        y2.mul_(-2.0).add_(y1, alpha=0.5).add_(zero, alpha=2.5)            # affine: bias counted once
        assert rel_err(y12, y2) < FP32_TOL
    del y2, y12, x2, zero
    # adjoint identities through autograd (the C-ABI backward)
    x1.requires_grad_(True)
    y = conv.hop(rel, x1, graph, precision="tf32x3")
    gy = torch.randn(n, f, device=DEV, generator=g)
    y.backward(gy)
    lin = _dot64(y.detach(), gy) - _dot64(conv.bias.detach().expand(n, f), gy)     # the part linear in x
    via_x = _dot64(x1.detach(), x1.grad)
    via_w = (_dot64(conv.weight.detach(), conv.weight.grad) + _dot64(conv.root.detach(), conv.root.grad))
    scale = (_dot64(y.detach(), y.detach()) * _dot64(gy, gy)) ** 0.5             # Cauchy-Schwarz scale of <y, g>
    assert abs(lin - via_x) < 1e-5 * scale, (lin, via_x, scale)
    assert abs(lin - via_w) < 1e-5 * scale, (lin, via_w, scale)
    gb = conv.bias.grad.double()
    assert float((gb - gy.double().sum(0)).abs().max()) < 1e-5 * float(gy.double().abs().sum(0).max())
