# times the tcgen05 kernels of one hop at 4M nodes (experiment helper, not part of the product)
import os, sys, torch
sys.path.insert(0, '.')
import mpgnn_b200
from mpgnn_b200 import _lib
lib = _lib.load()
dev = 'cuda'
n, e, r, f = 4_000_000, 40_000_000, 32, 128
gen = torch.Generator(device=dev).manual_seed(0)
ei = torch.randint(0, n, (2, e), device=dev, generator=gen); et = torch.randint(0, r, (e,), device=dev, generator=gen)
graph = mpgnn_b200.RelationGraph(ei, et, n, r); del ei, et
x = torch.randn(n, f, device=dev, generator=gen); gy = torch.randn(n, f, device=dev, generator=gen)
conv = mpgnn_b200.CustomRGCNConv(f, f, 1, flow='target_to_source', device=dev)
w, root, b = conv.weight.detach(), conv.root.detach(), conv.bias.detach()
h = torch.empty(n, f, device=dev); y = torch.empty(n, f, device=dev); gx = torch.empty(n, f, device=dev)
gw, gr, gb = torch.empty_like(w), torch.empty_like(root), torch.empty_like(b)
am = torch.empty(n, f // 32, dtype=torch.int32, device=dev)
ws = torch.empty(lib.mpgnn_hop_workspace_bytes(n, f, f), dtype=torch.uint8, device=dev)
st = _lib.current_stream()
ff = _lib.F_RELU | _lib.F_DROPOUT_SEED | _lib.F_TF32X3
lib.mpgnn_timing_reset(); lib.mpgnn_timing_enable(1)
for s in range(8):
    _lib.check(lib.mpgnn_hop_fwd(graph.handle, s % r, _lib.ptr(x), f, _lib.ptr(w), _lib.ptr(root), _lib.ptr(b), f, ff, 0.6, 1, s, None, _lib.ptr(h), _lib.ptr(y), _lib.ptr(am), _lib.ptr(ws), ws.numel(), st))
    _lib.check(lib.mpgnn_hop_bwd(graph.handle, s % r, _lib.ptr(x), _lib.ptr(h), None, _lib.ptr(am), _lib.ptr(gy), f, _lib.ptr(w), _lib.ptr(root), f, ff | _lib.F_NEED_GX, 0.6, _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(gr), _lib.ptr(gb), _lib.ptr(ws), ws.numel(), st))
torch.cuda.synchronize(); lib.mpgnn_timing_enable(0)
k = _lib.timing_collect()
print(os.environ.get('EXP_TAG', ''), ' '.join('%s %.3f' % (a, v[0] / v[1]) for a, v in k.items()), flush=True)
