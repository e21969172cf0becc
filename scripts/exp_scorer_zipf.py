# step-0 scorer (K5) on a power-law graph: hub destinations / sources (experiment helper)
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import mpgnn_b200
from mpgnn_b200 import search
n, r = 1_000_000, 100
g = torch.Generator().manual_seed(2)
deg = torch.randint(1, 6, (n,), generator=g)
rows = torch.repeat_interleave(torch.arange(n), deg); e = rows.numel()
for name in ("uniform", "zipf_dst", "zipf_src"):
    cols = torch.randint(0, n, (e,), generator=g)
    rr = rows
    if name != "uniform":
        u = torch.rand(e, generator=g)
        hub = (torch.exp(u * float(np.log(n))).long().clamp_(1, n) - 1)
        hub = torch.randperm(n, generator=g)[hub]
        if name == "zipf_dst": cols = hub
        else: rr = hub
    et = torch.randint(0, r, (e,), generator=g)
    lab = torch.randint(0, 2, (n,), generator=g).float()
    graph = mpgnn_b200.RelationGraph(torch.stack([rr, cols]), et, n, r, device="cuda")
    w0 = torch.rand(n, generator=g)
    search.run_scorer(graph, 0, w0, lab, epochs=5); torch.cuda.synchronize()
    t0 = time.time()
    for rel in range(10): search.run_scorer(graph, rel, w0, lab)
    torch.cuda.synchronize()
    print(name, "10 relations x 100 epochs: %.3f s" % (time.time() - t0), "max bucket", int(torch.bincount(rr * 0 + (cols if name == "zipf_dst" else rr)).max()))
