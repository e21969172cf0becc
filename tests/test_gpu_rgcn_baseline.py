"""The reference's all-relation RGCN comparison model (`Net`, model.py:132-151; main_rgcn.py:452-472) on the K1/K2
kernels and the library's dense kernels, against goldens recorded from the unmodified reference (rgcn_len3.npz)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

import mpgnn_b200
from mpgnn_b200 import rgcn_baseline as rb

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]
DEV = "cuda"


def _bag(fx):
    return mpgnn_b200.Data(**{k: fx[k] for k in ("x", "edge_index", "edge_type", "train_idx", "train_y", "val_idx",
                                                  "val_y", "test_idx", "test_y")}, num_nodes=fx["x"].size(0))


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def test_net_state_dict_forward_and_training_match_reference_golden(fx3):
    g = load_golden("rgcn_len3")
    r = fx3["num_relations"]
    torch.manual_seed(30)
    net = rb.Net(2, 64, r, 64, 2, 2, device="cpu")
    for k, v in _sd(g, "sd0.").items():
        assert torch.equal(net.state_dict()[k], v), k            # same draws in the same order as PyG's reset_parameters
    net.to(DEV)
    data = _bag(fx3)
    net.eval()
    with torch.no_grad():
        logp = net(fx3["x"].to(DEV), fx3["edge_index"].to(DEV), fx3["edge_type"].to(DEV))
    assert rel_err(logp, g["eval_logp"]) < 1e-5
    opt = torch.optim.Adam(net.parameters(), lr=0.01, weight_decay=0.0005)
    loss, _ = mpgnn_b200.mpgnn_train(net, opt, data)
    assert abs(loss - float(g["step_loss"])) < 1e-5 * abs(float(g["step_loss"]))
    for k, p in net.named_parameters():
        assert rel_err(p.grad, g["step_grad." + k]) < 5e-5, k
    ref = g["trace20"]
    f1t, f1v, _, lv = mpgnn_b200.mpgnn_validation(net, data, None)
    assert abs(float(lv) - ref[0, 1]) < 1e-4 * abs(ref[0, 1]) and abs(f1t - ref[0, 2]) < 2e-3 and abs(f1v - ref[0, 3]) < 2e-3
    for ep in range(1, 20):
        loss, cw = mpgnn_b200.mpgnn_train(net, opt, data)
        f1t, f1v, _, lv = mpgnn_b200.mpgnn_validation(net, data, cw)
        assert abs(loss - ref[ep, 0]) < 2e-4 * abs(ref[ep, 0]), ep
        assert abs(float(lv) - ref[ep, 1]) < 2e-4 * abs(ref[ep, 1]), ep
        assert abs(f1t - ref[ep, 2]) < 2e-3 and abs(f1v - ref[ep, 3]) < 2e-3, ep
    loss_t, f1_t = mpgnn_b200.mpgnn_test(net, data, None)
    assert abs(float(loss_t) - g["trace20_test"][0]) < 1e-3 * abs(g["trace20_test"][0])
    assert abs(f1_t - g["trace20_test"][1]) < 2e-3
    # conv2 applied twice (metapath_length = 3)
    torch.manual_seed(30)
    net3 = rb.Net(2, 64, r, 64, 2, 3)
    net3.eval()
    with torch.no_grad():
        assert rel_err(net3(fx3["x"].to(DEV), fx3["edge_index"].to(DEV), fx3["edge_type"].to(DEV)), g["eval_logp_len3"]) < 1e-5


def test_rgcn_conv_gradients_match_torch_autograd_of_the_same_formula():
    """out = sum_r mean_r(x) W_r + x root + b on a random multigraph (duplicate edges, an unused relation id), forward
    and every gradient against torch autograd of the dense formula in float64."""
    gen = torch.Generator().manual_seed(4)
    n, e, r, fi, fo = 700, 5000, 5, 24, 40
    ei = torch.randint(0, n, (2, e), generator=gen)
    et = torch.randint(0, r - 1, (e,), generator=gen)                # relation r-1 never occurs
    x = torch.randn(n, fi, generator=gen)
    gy = torch.randn(n, fo, generator=gen)
    torch.manual_seed(1)
    conv = rb.RGCNConv(fi, fo, r, flow="target_to_source", device=DEV)
    xi = x.to(DEV).requires_grad_(True)
    y = conv(xi, ei.to(DEV), et.to(DEV))
    y.backward(gy.to(DEV))
    xd = x.double().requires_grad_(True)
    w, root, b = (t.detach().cpu().double().requires_grad_(True) for t in (conv.weight, conv.root, conv.bias))
    out = xd @ root + b
    for rel in range(r):
        sel = et == rel
        a = torch.zeros(n, n, dtype=torch.float64)
        a.index_put_((ei[0][sel], ei[1][sel]), torch.ones(int(sel.sum()), dtype=torch.float64), accumulate=True)
        deg = a.sum(1).clamp(min=1)
        out = out + ((a @ xd) / deg[:, None]) @ w[rel]
    out.backward(gy.double())
    assert rel_err(y, out.detach()) < 1e-5
    for name, got, ref in (("x", xi.grad, xd.grad), ("weight", conv.weight.grad, w.grad), ("root", conv.root.grad, root.grad),
                           ("bias", conv.bias.grad, b.grad)):
        assert rel_err(got, ref) < 1e-5, name


def test_rgcn_training_call_learns_the_fixture(fx3):
    """main_rgcn.py:452-472 end to end (fewer epochs): the all-relation model also separates the fixture's classes."""
    torch.manual_seed(30)
    f1 = rb.mpgnn_parallel_multiple(_bag(fx3), 2, 64, fx3["num_relations"], 64, 2, 2, epochs=150, log=None)
    assert f1 > 0.9
