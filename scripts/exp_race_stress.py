# repeated tcgen05-vs-fp32 forward on shapes that exposed the raw-ring race; prints bad elements/rows per trial
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
    b = torch.zeros(f_out).cuda()
    graph = mpgnn_b200.RelationGraph(ei, et, n, 3, device='cuda')
    h32, y32 = t._fwd(graph, 1, x, w, root, b, 0, None)
    res = []
    for trial in range(8):
        htc, ytc = t._fwd(graph, 1, x, w, root, b, _lib.F_TF32X3, None)
        bad = (ytc - y32).abs() > 1e-3
        res.append((int(bad.sum()), int(bad.any(1).sum())))
    print((n, f_in, f_out), "bad (elems, rows) per trial:", res, flush=True)
for shp in [(50001, 96, 192), (50001, 32, 192), (400000, 64, 64), (50001, 64, 192), (50001, 96, 128), (400000, 128, 128)]:
    run(*shp)
