"""CPU-side checks of the host layer: the C-ABI library loads and exports every symbol the
header declares (no compute calls), the reference's constructor/parameter contract holds,
and the product path refuses to run without CUDA (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

import mpgnn_b200
from mpgnn_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpgnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpgnn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.PROTOTYPES) == declared  # the ctypes table mirrors the header
    assert lib.mpgnn_abi_version() >= 1


def test_state_dict_matches_reference_init():
    g = load_golden("model_len3")
    torch.manual_seed(30)
    m = mpgnn_b200.MPNetm(2, 64, 4, 64, 2, 1, [[1, 0]], device="cpu")
    sd = m.state_dict()
    keys = [k[4:] for k in g if k.startswith("sd0.")]
    assert sorted(sd.keys()) == sorted(keys)
    for k in keys:
        assert np.array_equal(sd[k].numpy(), g["sd0." + k]), k


def test_conv_contract():
    c = mpgnn_b200.CustomRGCNConv(4, 8, 3, flow="target_to_source", device="cpu")
    assert (c.in_channels, c.out_channels, c.num_relations) == (4, 8, 3)
    assert c.weight.shape == (4, 8) and c.root.shape == (4, 8) and c.bias.shape == (8,)
    assert c.comp is None and float(c.bias.abs().sum()) == 0.0
    assert "CustomRGCNConv(4, 8, num_relations=3)" == repr(c)
    with pytest.raises(ValueError):  # mp_rgcn_layer.py:106-108
        mpgnn_b200.CustomRGCNConv(4, 8, 3, num_bases=2, num_blocks=2, flow="target_to_source")
    with pytest.raises(NotImplementedError):
        mpgnn_b200.CustomRGCNConv(4, 8, 3, num_bases=2, flow="target_to_source")
    with pytest.raises(NotImplementedError):
        mpgnn_b200.CustomRGCNConv(4, 8, 3, flow="source_to_target")


def test_no_cpu_fallback():
    c = mpgnn_b200.CustomRGCNConv(2, 4, 1, flow="target_to_source", device="cpu")
    x = torch.zeros(3, 2)
    ei = torch.tensor([[0, 1], [1, 2]])
    et = torch.tensor([0, 0])
    with pytest.raises(RuntimeError):
        c(0, 0, x, ei, et)
    m = mpgnn_b200.MPNetm(2, 4, 1, 4, 2, 1, [[0]], device="cpu")
    with pytest.raises(RuntimeError):
        m(x, ei, et)


def test_masked_edge_index_matches_reference_golden():
    g = load_golden("layer_len3")
    fx = load_golden("fixture_len3")
    ei = torch.from_numpy(fx["edge_index"].astype(np.int64))
    et = torch.from_numpy(fx["edge_type"].astype(np.int64))
    for r in range(int(fx["num_relations"])):
        assert np.array_equal(mpgnn_b200.masked_edge_index(ei, et == r).numpy(), g["mei_r%d" % r])


def test_pack_mask_bits_layout():
    from mpgnn_b200.mp_rgcn_layer import pack_mask_bits
    rng = np.random.RandomState(0)
    for f in (8, 64, 13):
        m = rng.rand(5, f) > 0.5
        assert np.array_equal(pack_mask_bits(torch.from_numpy(m)).numpy(), np.packbits(m, axis=1))
