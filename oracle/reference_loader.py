"""Import the reference's own modules verbatim, in THIS container only.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may call this; it exists to (a) validate
``oracle/retrieval.py`` against the real ``src/search_engine.py`` and (b)
generate the golden vectors under ``tests/golden/`` (``oracle/make_golden.py``).

``search_engine.py`` imports two packages that are absent offline, and both are
needed only at import time / by the network branches we never call:
  * ``voyageai``  -- the annotation ``voyageai.Client`` at ``search_engine.py:16``
  * ``nltk`` (+ ``nltk.corpus.stopwords``, ``nltk.stem.WordNetLemmatizer``,
    ``nltk.tokenize.word_tokenize``, ``nltk.data.find``, ``nltk.download``) --
    imported and probed by ``processing/preprocess_bm25.py:6-23``
They are stubbed in ``sys.modules`` for the duration of the import.  No reference
source is copied: the files are executed from where they lie.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from contextlib import contextmanager

REFERENCE_ROOT = os.environ.get("ANR_REFERENCE_ROOT", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "search_engine.py"))


def _stub_modules():
    voyageai = types.ModuleType("voyageai")
    voyageai.Client = type("Client", (), {})

    nltk = types.ModuleType("nltk")
    nltk.data = types.SimpleNamespace(find=lambda *_a, **_k: True)
    nltk.download = lambda *_a, **_k: True
    corpus = types.ModuleType("nltk.corpus")
    corpus.stopwords = types.SimpleNamespace(words=lambda *_a, **_k: [])
    stem = types.ModuleType("nltk.stem")

    class WordNetLemmatizer:  # identity: the golden cases never tokenise text
        def lemmatize(self, tok):
            return tok

    stem.WordNetLemmatizer = WordNetLemmatizer
    tokenize = types.ModuleType("nltk.tokenize")
    tokenize.word_tokenize = lambda text: text.split()
    nltk.corpus, nltk.stem, nltk.tokenize = corpus, stem, tokenize
    return {
        "voyageai": voyageai, "nltk": nltk, "nltk.corpus": corpus,
        "nltk.stem": stem, "nltk.tokenize": tokenize,
    }


@contextmanager
def _reference_import_context():
    """sys.path / sys.modules arranged like ``cd src && python`` in the reference."""
    stubs = _stub_modules()
    shadowed = ["processing", "processing.preprocess_bm25", "search_engine",
                "database_manager", "config"]
    saved = {name: sys.modules.get(name) for name in list(stubs) + shadowed}
    saved_path = list(sys.path)
    try:
        for name in shadowed:
            sys.modules.pop(name, None)
        sys.modules.update(stubs)
        sys.path.insert(0, REFERENCE_SRC)
        yield
    finally:
        sys.path[:] = saved_path
        for name, mod in saved.items():
            if mod is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = mod


def _load(module_file: str, alias: str):
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REFERENCE_SRC, module_file))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load_reference():
    """Returns a namespace with the reference's SearchEngine, DatabaseManager, Config."""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    if "ns" not in _cache:
        with _reference_import_context():
            se = _load("search_engine.py", "_anr_ref_search_engine")
            dm = _load("database_manager.py", "_anr_ref_database_manager")
            cfg = _load("config.py", "_anr_ref_config")
            # the orchestrator imports the three modules above by their plain names
            sys.modules["search_engine"], sys.modules["database_manager"] = se, dm
            sys.modules["config"] = cfg
            qrr = _load("query_rag_retrieval.py", "_anr_ref_query_rag_retrieval")
        _cache["ns"] = types.SimpleNamespace(
            search_engine=se, database_manager=dm, config=cfg, query_rag_retrieval=qrr,
            SearchEngine=se.SearchEngine, DatabaseManager=dm.DatabaseManager, Config=cfg.Config,
            InfoSource=cfg.InfoSource, RetrievalEvaluationSystem=qrr.RetrievalEvaluationSystem,
        )
    return _cache["ns"]
