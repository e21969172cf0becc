# per-role cycle counters of CTA 0 of the tcgen05 projection kernel (experiment build only)
import ctypes, os, sys, torch
sys.path.insert(0, '.')
import mpgnn_b200
from mpgnn_b200 import _lib
lib = _lib.load()
dev = 'cuda'
n, e, r, f = 4_000_000, 40_000_000, 8, 128
gen = torch.Generator(device=dev).manual_seed(0)
ei = torch.randint(0, n, (2, e), device=dev, generator=gen); et = torch.randint(0, r, (e,), device=dev, generator=gen)
graph = mpgnn_b200.RelationGraph(ei, et, n, r); del ei, et
x = torch.randn(n, f, device=dev, generator=gen); gy = torch.randn(n, f, device=dev, generator=gen)
conv = mpgnn_b200.CustomRGCNConv(f, f, 1, flow='target_to_source', device=dev)
w, root, b = conv.weight.detach(), conv.root.detach(), conv.bias.detach()
h = torch.empty(n, f, device=dev); y = torch.empty(n, f, device=dev); gx = torch.empty(n, f, device=dev)
am = torch.empty(n, f // 32, dtype=torch.int32, device=dev)
gw, gr, gb = torch.empty_like(w), torch.empty_like(root), torch.empty_like(b)
ws = torch.empty(lib.mpgnn_hop_workspace_bytes(n, f, f), dtype=torch.uint8, device=dev)
st = _lib.current_stream()
ff = _lib.F_RELU | _lib.F_DROPOUT_SEED | _lib.F_TF32X3
raw = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * 64)()
names = {0: "conv g0 [total rfull empty st+arrive iters]", 1: "conv g1", 2: "epi w0 [total tfull ld - tiles]", 3: "tma [total rempty]", 4: "mma [total tempty full issue mma-only mma+commit]"}
def dump(tag):
    torch.cuda.synchronize(); raw.mpgnn_tc_debug_read(buf)
    print(tag)
    for role in range(5):
        print("   %-46s" % names[role], [int(buf[role * 8 + k]) for k in range(6)])
for s in range(3):
    _lib.check(lib.mpgnn_hop_fwd(graph.handle, s % r, _lib.ptr(x), f, _lib.ptr(w), _lib.ptr(root), _lib.ptr(b), f, ff, 0.6, 1, s, None, _lib.ptr(h), _lib.ptr(y), _lib.ptr(am), _lib.ptr(ws), ws.numel(), st))
dump("fwd exp=%s" % os.environ.get("MPGNN_TC_EXP", "0"))
for s in range(3):
    _lib.check(lib.mpgnn_hop_bwd(graph.handle, s % r, _lib.ptr(x), _lib.ptr(h), None, _lib.ptr(am), _lib.ptr(gy), f, _lib.ptr(w), _lib.ptr(root), f, ff | _lib.F_NEED_GX, 0.6, _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(gr), _lib.ptr(gb), _lib.ptr(ws), ws.numel(), st))
dump("dgrad exp=%s" % os.environ.get("MPGNN_TC_EXP", "0"))
