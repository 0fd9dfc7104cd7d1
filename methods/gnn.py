"""``methods.gnn`` as the reference scripts import it (``from methods.gnn import
GNN_nl`` -- gnnnet.py:5, gnnnet_copy.py:5, dampnet.py:6, dampnet_full.py:6;
``from methods import gnn`` -- train.py:12), served by the sm_100a kernels."""
from mft_b200.gnn import GNN_nl, Gconv, Wcompute, gmul  # noqa: F401
