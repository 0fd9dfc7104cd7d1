"""Drop-in overlay for the reference's ``methods`` package.

Put this repo's root on ``sys.path`` *before* the reference checkout: ``import
methods.gnn`` then resolves to the accelerated module below, while every other
submodule (``methods.gnnnet``, ``methods.meta_template``, ...) still resolves to
the reference's own files because the package path is extended over all
``methods`` directories on ``sys.path``.
"""
import pkgutil as _pkgutil

__path__ = _pkgutil.extend_path(__path__, __name__)
