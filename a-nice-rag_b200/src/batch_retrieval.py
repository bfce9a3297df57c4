"""Reference-style module name for the batched orchestrator: with this directory on sys.path,
``from batch_retrieval import retrieve_documents_batch`` resolves to the B200 implementation."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_impl = _importlib.import_module("a-nice-rag_b200.batch_retrieval")
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
