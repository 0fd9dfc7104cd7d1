"""Importable alias of the ``meta-fine-tuning_b200/`` package directory.

A hyphen cannot appear in a Python module name, so this shim points its
``__path__`` at the real package directory and executes that package's
``__init__``: ``import mft_b200`` / ``from mft_b200.gnn import GNN_nl``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "meta-fine-tuning_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
