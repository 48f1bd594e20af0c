"""Import shim: ``import cse_b200`` resolves to the package directory
``crowded-scenes-ensemble-classification_b200/`` (whose name is not a valid
Python identifier).  A module that defines ``__path__`` is a package as far as
the import system is concerned, so ``import cse_b200.graph`` etc. work.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                         "crowded-scenes-ensemble-classification_b200")
__path__ = [_PKG_DIR]
__file__ = _os.path.join(_PKG_DIR, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
