#!/bin/bash
# launch list of the candidate trainer at the C2 shape (eager launches, 3 epochs after warm-up)
mkdir -p gpurun_out
cat > /tmp/tr.py <<'PY'
import sys, torch
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
torch.manual_seed(30)
model = MPNetm(2, hidden, r, hidden, 2, 1, [[0, 1, 2]], device="cpu")
tr = CandidateTrainer(data, 2, hidden, 2, [0, 1, 2], dropout_p=0.6, max_epochs=40)
tr.load_state_dict(model.state_dict())
tr.run(3, use_graph=False); torch.cuda.synchronize()
PY
python /tmp/tr.py && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_trainer.csv python /tmp/tr.py > /dev/null 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows = [r for r in csv.reader(open("gpurun_out/launches_trainer.csv")) if len(r) > 5]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    k = r[ki].split("(")[0][:60]; d[k][0] += 1; d[k][1] += v
tot = sum(v[1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%-62s %5d launches %9.1f us total  %5.1f us each  %4.1f%%" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))
PY
