"""Import the reference's own classes from ``oracle/_ref`` (TEST INFRASTRUCTURE, see make_ref.py).

``load("reference")`` -> the reference exactly as it is (its own ``methods/gnn.py`` on ATen ops);
``load("overlay")``   -> the same reference files, but with this repo's root first on ``sys.path`` so that
                         ``methods.gnn`` resolves to the sm_100a kernels --
                         i.e. what a user gets who runs the reference's scripts with PYTHONPATH set.

Both variants can live in one process: each is imported into a private set of module objects (the
``methods`` / ``backbone`` / ``utils`` / ``configs`` / ``datasets`` entries of ``sys.modules`` are swapped
out around the import), so tests can build the reference's ``GnnNet`` twice and compare them.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ROOT = os.path.dirname(HERE)
_TOP = ("methods", "backbone", "utils", "configs", "datasets", "io_utils")
_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "methods", "gnnnet.py"))


def _is_ours(name):
    return name in _TOP or name.split(".")[0] in _TOP


def load(variant: str, cpu_shim: bool = False):
    """Returns a namespace with .gnn .gnnnet .gnnnet_copy .dampnet .dampnet_full .meta_template .backbone
    (modules).  ``cpu_shim``: make ``.cuda()`` the identity (the reference hard-codes it, gnnnet.py:40,69) --
    for the CPU legs of bench.py / CPU tests only."""
    assert variant in ("reference", "overlay")
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    key = (variant, cpu_shim)
    if key in _cache:
        return _cache[key]
    if cpu_shim:
        import torch
        import torch.nn as nn
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    saved_mods = {k: v for k, v in sys.modules.items() if _is_ours(k)}
    for k in saved_mods:
        del sys.modules[k]
    saved_path = list(sys.path)
    try:
        sys.path[:] = ([ROOT] if variant == "overlay" else []) + [REF_DIR] + \
            [p for p in saved_path if os.path.abspath(p or ".") not in (ROOT, REF_DIR)]
        importlib.invalidate_caches()
        ns = types.SimpleNamespace(variant=variant)
        ns.backbone = importlib.import_module("backbone")
        ns.gnn = importlib.import_module("methods.gnn")
        ns.meta_template = importlib.import_module("methods.meta_template")
        ns.gnnnet = importlib.import_module("methods.gnnnet")
        ns.gnnnet_copy = importlib.import_module("methods.gnnnet_copy")
        ns.dampnet = importlib.import_module("methods.dampnet")
        ns.dampnet_full = importlib.import_module("methods.dampnet_full")
        want = os.path.join(ROOT if variant == "overlay" else REF_DIR, "methods", "gnn.py")
        assert os.path.abspath(ns.gnn.__file__) == want, (ns.gnn.__file__, want)
        assert os.path.abspath(ns.gnnnet.__file__) == os.path.join(REF_DIR, "methods", "gnnnet.py")
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if _is_ours(k)]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        importlib.invalidate_caches()
    _cache[key] = ns
    return ns
