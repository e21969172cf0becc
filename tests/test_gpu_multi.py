"""Two-GPU run of the search fan-out over NCCL: every rank must derive the same decisions as a
single-process run (needs >= 2 visible GPUs; skipped otherwise)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]

WORKER = r'''
import json, os, sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch.distributed as dist
from conftest import fixture_as_torch
import mpgnn_b200
from mpgnn_b200 import search, main as m
world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fx = fixture_as_torch("fixture_len3")
data = mpgnn_b200.Data(x=fx["x"], edge_index=fx["edge_index"], edge_type=fx["edge_type"], labels=fx["labels"].unsqueeze(-1),
                       num_nodes=fx["x"].size(0), source_nodes_mask=[])
bag = mpgnn_b200.Data(**{k: fx[k] for k in ("x", "edge_index", "edge_type", "train_idx", "train_y", "val_idx", "val_y",
                                             "test_idx", "test_y")}, num_nodes=fx["x"].size(0))
ev = lambda meta: (torch.manual_seed(30), m.mpgnn_parallel_multiple(bag, 2, 64, 4, 64, 2, [meta], epochs=80))[1]
un = lambda metas: (torch.manual_seed(30), m.mpgnn_parallel_multiple_x(bag, 2, 64, 4, 64, 2, metas, True, epochs=80))[1]
res = search.greedy_search(data, bag, 2, 64, 4, 64, 2, "synthetic", comm=search.Comm(torch.device("cuda", local)),
                           eval_fn=ev, union_fn=un)
print("RESULT", os.environ.get("RANK", "0"), json.dumps(res), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
'''


def _run(nproc):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    code = WORKER % {"root": ROOT}
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), "-c", code] if nproc > 1 else [sys.executable, "-c", code]
    if nproc > 1:   # torchrun has no -c: go through a temp file
        path = os.path.join(ROOT, "gpurun_out", "_multi_worker.py")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        open(path, "w").write(code)
        cmd = cmd[:-2] + [path]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return {int(l.split(" ", 2)[1]): json.loads(l.split(" ", 2)[2]) for l in out.stdout.splitlines() if l.startswith("RESULT")}


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_search_two_gpus_matches_one_gpu():
    one = _run(1)[0]
    two = _run(2)
    assert set(two) == {0, 1}
    assert two[0] == two[1] == one          # identical decisions and scores on 1 and 2 GPUs
