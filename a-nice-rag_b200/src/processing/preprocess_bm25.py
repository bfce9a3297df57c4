"""Reference-named module (``from processing.preprocess_bm25 import preprocess_text``)."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_impl = _importlib.import_module("a-nice-rag_b200.processing.preprocess_bm25")
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
