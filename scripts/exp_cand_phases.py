# where does a candidate's wall time go? (experiment helper)
import sys, time, torch
sys.path.insert(0, '.')
import mpgnn_b200
from mpgnn_b200.main import CandidateTrainer, MPNetm
n, e, r, hidden = 100_000, 550_000, 20, 64
g = torch.Generator().manual_seed(1)
ei = torch.randint(0, n, (2, e), generator=g); et = torch.randint(0, r, (e,), generator=g)
x = torch.nn.functional.one_hot(torch.randint(0, 2, (n,), generator=g), 2).float()
y = torch.randint(0, 2, (n,), generator=g); perm = torch.randperm(n, generator=g)
n_te, n_va = n // 10, (n - n // 10) // 5
data = mpgnn_b200.Data(x=x, edge_index=ei, edge_type=et, num_nodes=n, test_idx=perm[:n_te], test_y=y[perm[:n_te]],
                       val_idx=perm[n_te:n_te + n_va], val_y=y[perm[n_te:n_te + n_va]], train_idx=perm[n_te + n_va:], train_y=y[perm[n_te + n_va:]])
def T():
    torch.cuda.synchronize(); return time.time()
for rep in range(3):
    t0 = T(); torch.manual_seed(30)
    model = MPNetm(2, hidden, r, hidden, 2, 1, [[0, 1, 2]], device="cpu"); t1 = T()
    tr = CandidateTrainer(data, 2, hidden, 2, [0, 1, 2], dropout_p=0.6, max_epochs=999); t2 = T()
    tr.load_state_dict(model.state_dict()); t3 = T()
    tr.run(1); t4 = T()
    tr.run(998); t5 = T()
    del tr; t6 = T()
    print("rep %d: model %.3f  create %.3f  load %.3f  first epoch (capture) %.3f  998 epochs %.3f  free %.3f" %
          (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5))
