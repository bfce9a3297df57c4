"""Drop-in ``DatabaseManager``: same contract as the reference's ``src/database_manager.py``
(:14-99), plus a device-resident copy of what it loads.

``load_embeddings_from_sql`` returns the same DataFrame (columns ``id, document, source,
embedding, url``; ``embedding`` an object column of per-row fp32 arrays) -- the rows are
views into ONE packed ``[N, D]`` matrix that is uploaded to HBM once, so searches never
re-stack it (the reference's ``np.stack`` per query, search_engine.py:80).
``load_bm25_from_pickle`` returns the same ``(bm25, sections, section_ids)`` tuple and
attaches the CSR inverted index built from the unpickled ``BM25Okapi`` attributes.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import pickle
import sqlite3
import sys
import threading
import types
from typing import Tuple

import numpy as np
import pandas as pd

from . import native, registry


def _no_device(exc: BaseException) -> bool:
    return isinstance(exc, native.AnrError) and exc.code == 3


def _alloc_matrix(n: int, d: int):
    """[n, d] fp32 host buffer; page-locked (faster, asynchronous H2D) when it is big and a GPU
    is present.  Returns (numpy view, owner object to keep alive)."""
    if n * d * 4 >= (64 << 20):
        try:
            import torch
            if torch.cuda.is_available():
                owner = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
                return owner.numpy(), owner
        except Exception:  # pragma: no cover - pinning is an optimisation only
            pass
    return np.empty((n, d), dtype=np.float32), None


class _AttrOnlyBM25Okapi:
    """Stand-in for ``rank_bm25.BM25Okapi`` when that package is absent at unpickle time.

    Real indices (written by src/processing/bm25_search.py:84-93) pickle an instance of
    that class by reference; only its attributes are needed here.  ``get_scores`` is
    served by the device index (fp32), there is no Python scoring loop behind it.
    """

    def get_scores(self, query):
        entry = registry.resolve_bm25(self)
        return entry.index.scores(entry.index.term_ids(query)).astype(np.float64)

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_anr_entry", None)   # device handles do not pickle
        return state


class _Document:
    """Stand-in for ``langchain.schema.document.Document`` (page_content + metadata)."""

    def __init__(self, page_content: str = "", metadata=None, **kwargs):
        self.page_content = page_content
        self.metadata = metadata if metadata is not None else {}
        self.__dict__.update(kwargs)

    def __setstate__(self, state):
        # pydantic-based Documents pickle {"__dict__": {...}, ...}; plain ones pickle the dict
        if isinstance(state, dict) and "__dict__" in state and isinstance(state["__dict__"], dict):
            state = state["__dict__"]
        self.__dict__.update(state)
        self.__dict__.setdefault("page_content", "")
        self.__dict__.setdefault("metadata", {})


def _install_unpickle_shims() -> None:
    """Make ``rank_bm25.BM25Okapi`` and langchain's ``Document`` resolvable for pickle.load
    when those packages are not installed (they are not, offline)."""
    try:
        import rank_bm25  # noqa: F401
    except ImportError:
        mod = types.ModuleType("rank_bm25")
        cls = type("BM25Okapi", (_AttrOnlyBM25Okapi,), {"__module__": "rank_bm25"})
        mod.BM25Okapi = cls
        sys.modules["rank_bm25"] = mod
    for name in ("langchain.schema.document", "langchain_core.documents.base"):
        try:
            __import__(name)
        except ImportError:
            parts = name.split(".")
            for i in range(1, len(parts) + 1):
                sub = ".".join(parts[:i])
                if sub not in sys.modules:
                    sys.modules[sub] = types.ModuleType(sub)
                    if i > 1:
                        setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[sub])
            doc_cls = type("Document", (_Document,), {"__module__": name})
            sys.modules[name].Document = doc_cls


class DatabaseManager:

    def __init__(self):
        self._embeddings_cache = {}
        self._bm25_cache = {}
        self._lock = threading.Lock()
        self.logger = logging.getLogger(__name__)

    def load_embeddings_from_sql(self, db_path: str, model_name: str = None) -> pd.DataFrame:
        """Load the ``chunks`` table into a DataFrame and pack + upload the embedding matrix."""
        cache_key = f"{db_path}_{model_name}" if model_name else db_path
        with self._lock:
            if cache_key in self._embeddings_cache:
                return self._embeddings_cache[cache_key]
        try:
            if not os.path.exists(db_path):
                raise FileNotFoundError(f"Database not found: {db_path}")
            conn = sqlite3.connect(db_path)
            cursor = conn.cursor()
            n_rows = cursor.execute("SELECT COUNT(*) FROM chunks").fetchone()[0]
            if not n_rows:
                self.logger.warning(f"No chunks found in {db_path}")
                return pd.DataFrame()
            # Fast path: the BLOBs are stepped through the SQLite C API by the library and copied
            # straight into ONE preallocated (pinned when a GPU is present) [N, D] buffer, while
            # this thread fetches the text columns (the library call runs without the GIL).
            # Tables it does not cover (NULL / ragged / non-BLOB embeddings, libsqlite3 absent)
            # are streamed row by row below, with the reference's skip-and-warn behaviour.
            ragged = None            # list of per-row arrays once widths disagree
            fast = self._load_uniform_table(conn, db_path, n_rows)
            if fast is not None:
                packed, keep, ids, documents, sources, urls = fast
                n_ok = len(ids)
            else:
                packed, keep, ragged, ids, documents, sources, urls, n_ok = \
                    self._load_row_by_row(cursor, n_rows)
            if ragged is None and packed is not None:
                packed = packed[:n_ok]
                embeddings = list(packed)            # N row views, like the per-row frombuffer
            else:                                    # searches will fail in np.stack like the
                embeddings = ragged or []            # reference does on such a table
            df = pd.DataFrame({
                "id": ids, "document": documents, "source": sources,
                "embedding": pd.Series(embeddings, dtype=object), "url": urls,
            })
            if packed is not None and len(df):
                entry = registry.register_frame(df, packed)
                entry.keepalive = keep               # the pinned allocation behind `packed`
                try:
                    entry.index()                    # upload now: queries must not pay for it
                except Exception as e:
                    if not _no_device(e):
                        raise
                    self.logger.warning(f"No B200 visible while loading {db_path}: "
                                        "the index will be uploaded by the first search")
            with self._lock:
                self._embeddings_cache[cache_key] = df
            return df
        except Exception as e:
            self.logger.error(f"Error loading embeddings from {db_path}: {e}")
            raise
        finally:
            if "conn" in locals():
                conn.close()

    def _load_uniform_table(self, conn, db_path: str, n_rows: int):
        """(packed, keep, ids, documents, sources, urls) when every row of ``chunks`` holds an
        embedding BLOB of one width (multiple of 4 bytes); None otherwise (the caller then takes
        the row-by-row path, which reproduces the reference's handling of odd rows)."""
        first = conn.execute("SELECT embedding FROM chunks LIMIT 1").fetchone()
        blob = first[0] if first else None
        if not isinstance(blob, bytes) or len(blob) == 0 or len(blob) % 4 != 0:
            return None
        d = len(blob) // 4
        packed, keep = _alloc_matrix(n_rows, d)
        rowids = np.empty(n_rows, dtype=np.int64)
        outcome = {}

        def read_blobs():
            got, uniform = C.c_int64(), C.c_int32()
            try:
                native.call("anr_sqlite_read_blobs", os.fsencode(db_path),
                            b"SELECT rowid, embedding FROM chunks", packed.ctypes.data, d * 4,
                            n_rows, rowids.ctypes.data, C.byref(got), C.byref(uniform))
                outcome["rows"], outcome["uniform"] = got.value, uniform.value
            except Exception as e:   # library without SQLite, unreadable file, ...
                outcome["error"] = e

        reader = threading.Thread(target=read_blobs, name="anr-sqlite-blobs")
        reader.start()
        try:
            rows = conn.execute("SELECT rowid, id, content, source, url FROM chunks").fetchall()
        except sqlite3.Error:
            rows = None     # e.g. a table without rowid / url: the row-by-row path reports it
        finally:
            reader.join()
        if rows is None:
            return None
        if "error" in outcome:
            self.logger.info(f"Bulk BLOB loader unavailable ({outcome['error']}); reading row by row")
            return None
        if not outcome.get("uniform") or outcome.get("rows") != len(rows) or len(rows) != n_rows:
            return None
        if not rows:
            return None
        rid, ids, documents, sources, urls = (list(c) for c in zip(*rows))
        # both statements scan the table in rowid order; make sure they saw the same rows
        if not np.array_equal(rowids, np.asarray(rid, dtype=np.int64)):
            return None
        return packed, keep, ids, documents, sources, urls

    def _load_row_by_row(self, cursor, n_rows: int):
        """The reference's loop (database_manager.py:46-61) streaming into one matrix: rows whose
        BLOB cannot be read as fp32 are skipped with a warning; D is fixed by the first decodable
        row; a table with ragged widths keeps per-row arrays like the reference's frame."""
        cursor.execute("SELECT id, content, source, embedding, url FROM chunks")
        ids, documents, sources, urls = [], [], [], []
        packed = keep = None
        ragged = None
        n_ok = 0
        while True:
            rows = cursor.fetchmany(16384)
            if not rows:
                break
            for cid, content, source, blob, url in rows:
                try:
                    view = memoryview(blob)
                    if len(view) % 4 != 0:
                        raise ValueError("buffer size must be a multiple of element size")
                except (ValueError, TypeError) as e:
                    self.logger.warning(f"Skipping invalid row {cid}: {e}")
                    continue
                row = np.frombuffer(view, dtype=np.float32)
                if packed is None and ragged is None:
                    packed, keep = _alloc_matrix(n_rows, row.shape[0])
                if ragged is None and row.shape[0] != packed.shape[1]:
                    ragged = [np.array(packed[i]) for i in range(n_ok)]
                    packed = keep = None
                if ragged is None:
                    packed[n_ok] = row
                else:
                    ragged.append(row)
                ids.append(cid)
                documents.append(content)
                sources.append(source)
                urls.append(url)
                n_ok += 1
        return packed, keep, ragged, ids, documents, sources, urls, n_ok

    def load_bm25_from_pickle(self, filepath: str) -> Tuple:
        """Unpickle ``{bm25, sections, section_ids}`` and build the device CSR index."""
        with self._lock:
            if filepath in self._bm25_cache:
                return self._bm25_cache[filepath]
        try:
            if not os.path.exists(filepath):
                raise FileNotFoundError(f"BM25 index not found: {filepath}")
            _install_unpickle_shims()
            with open(filepath, "rb") as f:
                data = pickle.load(f)
            result = (data["bm25"], data["sections"], data["section_ids"])
            try:
                registry.resolve_bm25(result[0], cache_for=filepath)   # CSR (cached on disk) + upload
            except Exception as e:
                if not _no_device(e):
                    raise
                self.logger.warning(f"No B200 visible while loading {filepath}: "
                                    "the index will be built by the first search")
            with self._lock:
                self._bm25_cache[filepath] = result
            return result
        except Exception as e:
            self.logger.error(f"Error loading BM25 index from {filepath}: {e}")
            raise
