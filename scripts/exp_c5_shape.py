"""Shape of the greedy search tree on the configs[4]-style graph at a reduced node count (GPU box)."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpgnn_b200
from mpgnn_b200 import synthetic, search
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 19
presets = ([0] + list(range(4, 51)), [1, 2], [3, 51], list(range(52, 100)))      # red-red 48, red-blue 2, blue-red 2, blue-blue 48
sg = synthetic.generate(n, deg, "red-blue-red-blue", 0, 0, seed=5, presets=presets)
x, ei, et, y = sg.tensors()
print("nodes", n, "edges", ei.size(1), "positives", int(y.sum()), "planted", sg.planted_relations, flush=True)
perm = torch.randperm(n, generator=torch.Generator().manual_seed(5))
n_te, n_va = n // 10, (n - n // 10) // 5
dm = mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, num_nodes=n, test_idx=perm[:n_te], test_y=y[perm[:n_te]],
                     val_idx=perm[n_te:n_te + n_va], val_y=y[perm[n_te:n_te + n_va]], train_idx=perm[n_te + n_va:],
                     train_y=y[perm[n_te + n_va:]])
data = mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, labels=y.unsqueeze(-1), num_nodes=n, source_nodes_mask=[])
tm = {}
t0 = time.time()
res = search.greedy_search(data, dm, 2, 64, 100, 64, 2, "synthetic", max_depth=3, epochs=int(os.environ.get("EPOCHS", "999")),
                           timings=tm, log=lambda s: print(s[:300], flush=True))
print("total %.1f s" % (time.time() - t0), tm, "final", res["final_meta"], res["test_f1"], "candidates", len(res["candidates"]))
