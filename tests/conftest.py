import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def fixture_as_torch(name="fixture_len3"):
    g = load_golden(name)
    return dict(
        x=torch.from_numpy(g["x"]), edge_index=torch.from_numpy(g["edge_index"].astype(np.int64)),
        edge_type=torch.from_numpy(g["edge_type"].astype(np.int64)), num_relations=int(g["num_relations"]),
        labels=torch.from_numpy(g["labels"].astype(np.int64)),
        train_idx=g["train_idx"].astype(np.int64).tolist(), train_y=torch.from_numpy(g["train_y"].astype(np.int64)),
        val_idx=g["val_idx"].astype(np.int64).tolist(), val_y=torch.from_numpy(g["val_y"].astype(np.int64)),
        test_idx=g["test_idx"].astype(np.int64).tolist(), test_y=torch.from_numpy(g["test_y"].astype(np.int64)))


def write_fixture_files(folder, name="fixture_len3"):
    """node.dat / link.dat / label.dat in the reference's TSV formats (`id \\t f0 \\t f1`, `src \\t rel \\t dst`,
    `id \\t label`), rewritten from a committed golden of the reference's own fixture folder (the folder itself does
    not exist on the GPU box)."""
    g = load_golden(name)
    x, ei, et, lab = g["x"], g["edge_index"], g["edge_type"], g["labels"]
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "node.dat"), "w") as f:
        f.write("".join("%d\t%s\n" % (i, "\t".join("%d" % v for v in row)) for i, row in enumerate(x)))
    with open(os.path.join(folder, "link.dat"), "w") as f:
        f.write("".join("%d\t%d\t%d\n" % (s, r, d) for s, r, d in zip(ei[0], et, ei[1])))
    with open(os.path.join(folder, "label.dat"), "w") as f:
        f.write("".join("%d\t%d\n" % (i, l) for i, l in enumerate(lab)))
    return folder


def rel_err(a, b):
    """Normalised max error: max|a-b| / max(|b|) -- the 'relative' of the 1e-5 fp32 bar."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


@pytest.fixture(scope="session")
def fx3():
    return fixture_as_torch("fixture_len3")


@pytest.fixture(scope="session")
def fx4():
    return fixture_as_torch("fixture_len4")
