import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import mpgnn_b200
from mpgnn_b200 import _lib
import test_gpu_tcgen05 as t
def run(n, f_in, f_out):
    ei, et = t._graph(n, 6 * n, 3, seed=n)
    gen = torch.Generator().manual_seed(n + 1)
    x = torch.randn(n, f_in, generator=gen).cuda()
    w = (torch.randn(f_in, f_out, generator=gen) * 0.1).cuda(); root = (torch.randn(f_in, f_out, generator=gen) * 0.1).cuda()
    b = (torch.randn(f_out, generator=gen) * 0.1).cuda()
    graph = mpgnn_b200.RelationGraph(ei, et, n, 3, device='cuda')
    h32, y32 = t._fwd(graph, 1, x, w, root, b, _lib.F_RELU, None)
    res = []
    for trial in range(3):
        htc, ytc = t._fwd(graph, 1, x, w, root, b, _lib.F_RELU | _lib.F_TF32X3, None)
        d = (ytc - y32).abs()
        res.append((float(d.max()), int((d > 1e-3).sum())))
    print((n, f_in, f_out), "max err / bad elems per trial:", res, flush=True)
for shp in [(300, 128, 128), (5000, 128, 128), (50001, 128, 128), (50001, 96, 128), (200000, 128, 256)]:
    run(*shp)
