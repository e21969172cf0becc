# candidate epoch on the random C2-shaped graph vs the generated one (experiment helper)
import sys, time, torch
sys.path.insert(0, '.')
import mpgnn_b200
from mpgnn_b200 import _lib, synthetic
from mpgnn_b200.main import CandidateTrainer, MPNetm
lib = _lib.load()
hidden = 64


def bag(x, ei, et, y, n):
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1))
    n_te, n_va = n // 10, (n - n // 10) // 5
    return mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, num_nodes=n, test_idx=perm[:n_te], test_y=y[perm[:n_te]],
                           val_idx=perm[n_te:n_te + n_va], val_y=y[perm[n_te:n_te + n_va]], train_idx=perm[n_te + n_va:],
                           train_y=y[perm[n_te + n_va:]])


n, e, r = 100_000, 550_000, 20
g = torch.Generator().manual_seed(1)
ei = torch.randint(0, n, (2, e), generator=g); et = torch.randint(0, r, (e,), generator=g)
x = torch.nn.functional.one_hot(torch.randint(0, 2, (n,), generator=g), 2).float()
y = torch.randint(0, 2, (n,), generator=g)
cases = [("random", bag(x, ei, et, y, n), r, [0, 1, 2])]
sg = synthetic.generate(100_000, 10, "red-blue-red-blue", 0, 2, seed=1)
sx, sei, set_, sy = sg.tensors()
cases.append(("generated", bag(sx, sei, set_, sy, n), int(set_.max()) + 1, sg.planted_relations))
cases.append(("generated-other", cases[1][1], cases[1][2], [0, 1, 2]))
for name, data, nrel, meta in cases:
    torch.manual_seed(30)
    model = MPNetm(2, hidden, nrel, hidden, 2, 1, [meta], device="cpu")
    tr = CandidateTrainer(data, 2, hidden, 2, meta, dropout_p=0.6, max_epochs=400, precision="tf32x3")
    tr.load_state_dict(model.state_dict())
    tr.run(20); torch.cuda.synchronize()
    t0 = time.time(); tr.run(200); torch.cuda.synchronize(); dt = time.time() - t0
    print(name, meta, "graph replay: %.3f ms/epoch" % (dt / 200 * 1e3), "val f1", tr.last_val_f1, flush=True)
    lib.mpgnn_timing_reset(); lib.mpgnn_timing_enable(1)
    tr.run(20, use_graph=False); torch.cuda.synchronize(); lib.mpgnn_timing_enable(0)
    k = _lib.timing_collect()
    tot = sum(v[0] for v in k.values()) / 20
    print("   timed kernels %.3f ms/epoch:" % tot, " ".join("%s %.3f" % (a, v[0] / 20) for a, v in sorted(k.items(), key=lambda kv: -kv[1][0])), flush=True)
    torch.manual_seed(30)
    t0 = time.time(); f1 = mpgnn_b200.mpgnn_parallel_multiple(data, 2, hidden, nrel, hidden, 2, [meta], epochs=999); torch.cuda.synchronize()
    print("   mpgnn_parallel_multiple 999 epochs: %.3f s, f1 %.4f" % (time.time() - t0, f1), flush=True)
