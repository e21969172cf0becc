"""Experiment: single-operand weight gradient (x^T g, 128 x 128) through the C-ABI against a float64 product, at row
counts around the chunk / CTA boundaries.  MPGNN_WGRAD_EXP selects the kernel variant (see wgrad_tcgen05.cu)."""
import sys

import torch

sys.path.insert(0, ".")
import mpgnn_b200
from mpgnn_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
worst = 0.0
for m in (1, 31, 32, 33, 4735, 148 * 32, 148 * 32 + 1, 100003, 1_000_000):
    g = torch.Generator(device=dev).manual_seed(m)
    a = torch.randn(m, 128, device=dev, generator=g)
    b = torch.randn(m, 128, device=dev, generator=g)
    out = torch.full((128, 128), float("nan"), device=dev)
    cs = torch.full((128,), float("nan"), device=dev)
    wsb = lib.mpgnn_gemm_workspace_bytes(m, 128, 128)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    for rep in range(2):
        _lib.check(lib.mpgnn_gemm_tn(_lib.ptr(a), 128, m, 128, _lib.ptr(b), 128, 128, _lib.ptr(out), 128, _lib.ptr(cs),
                                     _lib.ptr(ws), wsb, None))
        torch.cuda.synchronize()
        ref = a.double().t() @ b.double()
        e1 = float((out.double() - ref).abs().max() / ref.abs().max())
        e2 = float((cs.double() - b.double().sum(0)).abs().max() / b.double().sum(0).abs().max())
        worst = max(worst, e1, e2)
        print("m=%d rep %d: rel err out %.2e colsum %.2e" % (m, rep, e1, e2), flush=True)
assert worst < 1e-5, worst
print("OK worst", worst)
