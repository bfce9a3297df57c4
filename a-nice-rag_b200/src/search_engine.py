"""Reference-named module: put this directory on sys.path in place of the reference's
``src/`` and ``from search_engine import ...`` resolves to the B200 implementation."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_impl = _importlib.import_module("a-nice-rag_b200.search_engine")
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
