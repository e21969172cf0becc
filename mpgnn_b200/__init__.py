"""Importable alias of the package directory `mpgnn-metapath-graph-neural-network_b200/`
(a hyphenated name cannot be imported directly): `import mpgnn_b200` resolves its
submodules (`mp_rgcn_layer`, `model`, `main`, `graph`, `_lib`, `_build`) from there."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "mpgnn-metapath-graph-neural-network_b200")
__path__.append(_PKG_DIR)

from . import _lib  # noqa: E402,F401
from .graph import RelationGraph, graph_for, clear_cache  # noqa: E402,F401
from .mp_rgcn_layer import CustomRGCNConv, masked_edge_index  # noqa: E402,F401
from .model import MPNetm  # noqa: E402,F401
from .main import (Data, mpgnn_train, mpgnn_validation, mpgnn_test, mpgnn_parallel_multiple,  # noqa: E402,F401
                   mpgnn_parallel_multiple_x, mpgnn_parallel_multiple_batch, device_macro_f1, CandidateTrainer)
from .search import (score_relation_parallel, node_types_and_connected_relations, create_edge_dictionary,  # noqa: E402,F401
                     greedy_search, Comm, run_scorer)
from . import rgcn_baseline  # noqa: E402,F401  (Net / RGCNConv: the reference's all-relation comparison model)
