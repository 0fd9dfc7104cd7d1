"""B200-native episodic GNN few-shot head (drop-in for methods/gnn.py of
johncai117/Meta-Fine-Tuning).

The directory name carries a hyphen, so the package is imported through the
``mft_b200`` shim at the repo root (``import mft_b200``); ``methods/gnn.py``
re-exports the module classes under the reference's own import path.
"""
from .gnn import GNN_nl, Gconv, Wcompute, gmul, set_precision, get_precision  # noqa: F401
from .episode import GnnHead, support_label, build_graphs, select_scores, query_labels  # noqa: F401
from .graphs import GraphedStep  # noqa: F401
from ._lib import lib_path, load_library, LibraryMissing  # noqa: F401

__all__ = ["GNN_nl", "Gconv", "Wcompute", "gmul", "set_precision", "get_precision",
           "lib_path", "load_library", "LibraryMissing"]
