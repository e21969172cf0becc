# does the epoch time of a candidate drift over training? (experiment helper)
import sys, time, torch
sys.path.insert(0, '.')
import mpgnn_b200
from mpgnn_b200 import _lib, synthetic
from mpgnn_b200.main import CandidateTrainer, MPNetm
lib = _lib.load()
hidden = 64
sg = synthetic.generate(100_000, 10, "red-blue-red-blue", 0, 2, seed=1)
x, ei, et, y = sg.tensors()
n = sg.num_nodes
perm = torch.randperm(n, generator=torch.Generator().manual_seed(1))
n_te, n_va = n // 10, (n - n // 10) // 5
data = mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, num_nodes=n, test_idx=perm[:n_te], test_y=y[perm[:n_te]],
                       val_idx=perm[n_te:n_te + n_va], val_y=y[perm[n_te:n_te + n_va]], train_idx=perm[n_te + n_va:],
                       train_y=y[perm[n_te + n_va:]])
nrel = int(et.max()) + 1
for meta in (sg.planted_relations, [0, 1, 2]):
    torch.manual_seed(30)
    model = MPNetm(2, hidden, nrel, hidden, 2, 1, [meta], device="cpu")
    t0 = time.time()
    tr = CandidateTrainer(data, 2, hidden, 2, meta, dropout_p=0.6, max_epochs=1000)
    tr.load_state_dict(model.state_dict())
    torch.cuda.synchronize()
    print(meta, "create %.3f s" % (time.time() - t0), flush=True)
    chunks = []
    for c in range(9):
        t0 = time.time(); tr.run(100); torch.cuda.synchronize(); chunks.append(time.time() - t0)
    print("   100-epoch chunks (s):", " ".join("%.3f" % c for c in chunks), "f1", tr.last_val_f1, flush=True)
    lib.mpgnn_timing_reset(); lib.mpgnn_timing_enable(1)
    tr.run(20, use_graph=False); torch.cuda.synchronize(); lib.mpgnn_timing_enable(0)
    k = _lib.timing_collect()
    print("   late epochs, timed kernels:", " ".join("%s %.3f" % (a, v[0] / 20) for a, v in sorted(k.items(), key=lambda kv: -kv[1][0])), flush=True)
    torch.manual_seed(30)
    t0 = time.time(); f1 = mpgnn_b200.mpgnn_parallel_multiple(data, 2, hidden, nrel, hidden, 2, [meta], epochs=999); torch.cuda.synchronize()
    print("   mpgnn_parallel_multiple 999 epochs: %.3f s" % (time.time() - t0), flush=True)
