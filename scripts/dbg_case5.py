import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import mpgnn_b200
from mpgnn_b200 import _lib
import test_gpu_tcgen05 as t
n, f_in, f_out = 400000, 64, 64
ei, et = t._graph(n, 6 * n, 3, seed=n)
gen = torch.Generator().manual_seed(n + 1)
x = torch.randint(-8, 9, (n, f_in), generator=gen).float().cuda()
b = torch.zeros(f_out).cuda()
graph = mpgnn_b200.RelationGraph(ei, et, n, 3, device='cuda')
eye = torch.eye(f_in).cuda()
for mode in ("x", "h"):
    w = eye * (1.0 if mode == "h" else 0.0); root = eye * (1.0 if mode == "x" else 0.0)
    h32, y32 = t._fwd(graph, 1, x, w, root, b, 0, None)
    src = x if mode == "x" else h32
    for trial in range(4):
        htc, ytc = t._fwd(graph, 1, x, w, root, b, _lib.F_TF32X3, None)
        bad = (ytc - y32).abs() > 1e-3
        rows = bad.any(1).nonzero().flatten()
        tiles = (rows // 128)
        print("mode", mode, "trial", trial, "bad rows", rows.numel(), "ti hist", torch.bincount(tiles // 148).tolist()[:8],
              "chunk hist (cols 0-31, 32-63)", [int(bad[:, :32].sum()), int(bad[:, 32:].sum())])
        for r in rows[:3].tolist():
            cols = bad[r].nonzero().flatten().tolist()
            print("   row %d tile %d lane %d badcols %s" % (r, r // 128, r % 128, cols))
            print("      got ", [round(v, 3) for v in ytc[r, cols[:8]].tolist()], " want", [round(v, 3) for v in y32[r, cols[:8]].tolist()])
