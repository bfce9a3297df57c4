"""Finding the device-resident index that belongs to a caller's Python object.

The reference's callers keep the loaded ``DataFrame`` / ``BM25Okapi`` and hand it back
into every search call (SURVEY.md 8(b) "Ownership"), so the HBM copy must be
discoverable from that argument:
  * DataFrames carry ``df.attrs["anr_dense_key"]``; pandas propagates ``attrs`` to frames
    derived by masking/copying, which we recognise as row subsets of the original
    (their index labels are the original row positions);
  * BM25 objects get the index as an attribute (fallback: a registry keyed by ``id``).
A frame / object we have never seen is packed and uploaded on first use and cached
for as long as the Python object lives.
"""
from __future__ import annotations

import itertools
import logging
import os
import threading
import weakref
from typing import Dict, Optional, Tuple

import numpy as np
import pandas as pd

from . import engine

logger = logging.getLogger(__name__)
ATTR_KEY = "anr_dense_key"
_lock = threading.Lock()
_next_key = itertools.count(1)


class DenseEntry:
    """Host-side companion of one DenseIndex: packed rows, ids, sources, cached masks."""

    def __init__(self, packed: np.ndarray, ids: np.ndarray, sources: np.ndarray):
        self.key = next(_next_key)
        self.packed = packed            # [n, d] float32, C-contiguous (row i = df.iloc[i])
        self.ids = ids                  # object array of chunk ids, for subset validation
        self.sources = sources          # object array of `source` strings, for filters
        self.n, self.d = packed.shape
        # (row, data pointer) of a few `embedding` cells of the frame this entry was built from:
        # a frame that carries the same ids but OTHER vectors (df.copy() followed by a new or
        # renormalised embedding column; the column replaced on the loaded frame itself) must not
        # resolve to this device matrix -- the reference re-stacks df["embedding"] on every call
        # (search_engine.py:80).  None = rows are views into `packed` (the loader's frames).
        self.probe_ptrs: Optional[Dict[int, int]] = None
        self._index: Optional[engine.DenseIndex] = None
        self._masks: Dict[str, Tuple[np.ndarray, object, int]] = {}
        self._build_lock = threading.Lock()

    def index(self) -> engine.DenseIndex:
        with self._build_lock:
            if self._index is None:
                self._index = engine.DenseIndex(self.packed)
                # bf16 shadow copy for the tensor-core nomination pass (half the HBM bytes per
                # search, +50 % device memory, identical results); ANR_DENSE_SHADOW=0 turns it off.
                # Only indices large enough for that pass (>= ~76k rows, d % 64 == 0) ever build it.
                if os.environ.get("ANR_DENSE_SHADOW", "1") != "0" and self.d % 64 == 0:
                    self._index.set_shadow(True)
            return self._index

    def filter_mask(self, filename_type_filter: str):
        """(bool[n], device words or None, eligible count), cached per filter string."""
        hit = self._masks.get(filename_type_filter)
        if hit is None:
            mask = engine.prefix_mask(self.sources, filename_type_filter, frame_semantics=True)
            hit = (mask, device_words(engine.pack_mask(mask)), int(mask.sum()))
            self._masks[filename_type_filter] = hit
        return hit


def device_words(words: np.ndarray):
    """uint32 mask words -> resident device tensor (torch is the allocator here)."""
    import torch
    if not torch.cuda.is_available():
        return None
    return torch.from_numpy(words.view(np.int32)).to(f"cuda:{engine.current_device()}")


_dense: Dict[int, DenseEntry] = {}


def register_frame(df: pd.DataFrame, packed: np.ndarray) -> DenseEntry:
    entry = DenseEntry(packed, df["id"].to_numpy(dtype=object), df["source"].to_numpy(dtype=object))
    with _lock:
        _dense[entry.key] = entry
    df.attrs[ATTR_KEY] = entry.key
    weakref.finalize(df, _dense.pop, entry.key, None)
    return entry


def _cell_ptr(cell) -> int:
    """Address of the first element of one `embedding` cell (-1 when it is not an ndarray)."""
    return cell.__array_interface__["data"][0] if isinstance(cell, np.ndarray) else -1


def _cells(col):
    """The values of a column as something indexable by position without a copy (the column's own
    ndarray or extension array): a Series.iat call costs ~10 us, and the probes below run on every
    query.  (Never to_numpy(): on an Arrow-backed string column that converts all N rows.)"""
    values = getattr(col, "_values", None)      # ndarray or ExtensionArray: both take values[i]
    return values if values is not None and hasattr(values, "__getitem__") else col.iat


_probe_cache: Dict[int, np.ndarray] = {}


def _probe_positions(n: int) -> np.ndarray:
    hit = _probe_cache.get(n)
    if hit is None:
        if len(_probe_cache) > 64:
            _probe_cache.clear()
        hit = _probe_cache[n] = np.unique(np.linspace(0, n - 1, num=min(n, 8)).astype(np.int64))
    return hit


def _same_vectors(entry: DenseEntry, emb_col, positions, rows) -> bool:
    """Do the probed `embedding` cells of a frame still hold the memory the entry was packed
    from?  O(1): a handful of pointer compares per query.  (Writing INTO the loader's arrays in
    place is not detectable this way; DenseIndex.invalidate / a fresh load is the way to do that.)"""
    row_bytes = entry.packed.strides[0]
    base = entry.packed.ctypes.data
    cells = _cells(emb_col)
    for p, r in zip(positions, rows):
        cell = cells[int(p)]
        want = base + int(r) * row_bytes if entry.probe_ptrs is None else entry.probe_ptrs.get(int(r))
        if want is None:
            continue                      # this row was not probed at registration
        if _cell_ptr(cell) != want or cell.shape != (entry.d,) or cell.dtype != np.float32:
            return False
    return True


def resolve_frame(df: pd.DataFrame) -> Tuple[DenseEntry, Optional[np.ndarray]]:
    """-> (entry, None) when ``df`` is a registered frame, (entry, row positions) when it is
    a row subset derived from one; an unknown frame is packed, uploaded and remembered."""
    entry = lookup_identity(df)
    if entry is not None:
        n = len(df)
        if n == entry.n and "embedding" in df.columns and \
                (n == 0 or _same_vectors(entry, df["embedding"], sorted(entry.probe_ptrs or {}),
                                         sorted(entry.probe_ptrs or {}))):
            return entry, None
        _forget_identity(id(df), entry.key)     # same object, new vectors: pack it again
    entry = _dense.get(df.attrs.get(ATTR_KEY, -1))
    if entry is not None and "id" in df.columns and "embedding" in df.columns:
        # O(1) checks only: this runs on every query (a full pass over a 1M-row id column
        # costs more than the search itself)
        n = len(df)
        labels = df.index
        emb_col = df["embedding"]          # (one column fetch costs ~60 us in pandas 3: no more
        if n == entry.n and isinstance(labels, pd.RangeIndex) and labels.start == 0 \
                and labels.step == 1:      #  than needed on the per-query path)
            # the whole frame in load order: the probed cells ARE rows 0, .., n - 1 of the packed
            # matrix (first and last included); the result frame is taken from df itself, so its
            # other columns need no check
            probe = _probe_positions(n) if n else np.zeros(0, dtype=np.int64)
            if n == 0 or _same_vectors(entry, emb_col, probe, probe):
                return entry, None
        id_col = df["id"]
        rows = np.asarray(labels)
        if n and rows.dtype.kind in "iu" and rows.min() >= 0 and rows.max() < entry.n \
                and labels.is_unique:
            probe = _probe_positions(n)
            ids = _cells(id_col)
            if all(ids[int(p)] == entry.ids[rows[p]] for p in probe) and \
                    _same_vectors(entry, emb_col, probe, rows[probe]):
                return entry, rows.astype(np.int64)
    # unknown frame: pack and upload it (np.stack raises for ragged rows, like the reference)
    packed = np.ascontiguousarray(np.stack(df["embedding"].values), dtype=np.float32)
    if packed.ndim != 2:
        raise ValueError("embedding column does not stack to a matrix")
    n = len(df)
    entry = DenseEntry(
        packed,
        df["id"].to_numpy(dtype=object) if "id" in df.columns else np.arange(n).astype(object),
        df["source"].to_numpy(dtype=object) if "source" in df.columns
        else np.full(n, None, dtype=object))
    # packed is a copy (np.stack): remember where the probed cells lived instead
    emb_col = df["embedding"]
    entry.probe_ptrs = {int(p): _cell_ptr(emb_col.iat[int(p)]) for p in _probe_positions(n)} if n else {}
    with _lock:
        _dense[entry.key] = entry
    # an unknown frame may carry arbitrary index labels: remember it by identity only
    _by_identity[id(df)] = (weakref.ref(df), entry)
    weakref.finalize(df, _forget_identity, id(df), entry.key)
    return entry, None


_by_identity: Dict[int, Tuple[weakref.ref, DenseEntry]] = {}


def _forget_identity(obj_id: int, key: int) -> None:
    _by_identity.pop(obj_id, None)
    _dense.pop(key, None)


def lookup_identity(df: pd.DataFrame) -> Optional[DenseEntry]:
    hit = _by_identity.get(id(df))
    if hit is not None and hit[0]() is df:
        return hit[1]
    return None


# ---------------------------------------------------------------------------------
CSR_CACHE_SUFFIX = ".anr_csr.npz"


def _csr_cache_load(pickle_path: str):
    """CSR arrays saved next to the BM25 pickle by an earlier load, if still valid (same pickle
    size and mtime).  The Python inversion of a real index takes minutes at 1M documents; the
    cache makes it a one-off."""
    import os
    path = pickle_path + CSR_CACHE_SUFFIX
    try:
        st = os.stat(pickle_path)
        with np.load(path, allow_pickle=True) as z:
            if int(z["pickle_size"]) != st.st_size or int(z["pickle_mtime_ns"]) != st.st_mtime_ns:
                return None
            return {k: z[k] for k in ("vocab", "term_ptr", "post_doc", "post_tf", "doc_len", "idf",
                                      "k1", "b", "avgdl")}
    except (OSError, KeyError, ValueError):
        return None


def _csr_cache_save(pickle_path: str, vocab, term_ptr, post_doc, post_tf, doc_len, idf, k1, b,
                    avgdl) -> None:
    import os
    try:
        st = os.stat(pickle_path)
        tmp = pickle_path + CSR_CACHE_SUFFIX + ".tmp.npz"
        np.savez(tmp, vocab=np.array(list(vocab.keys()), dtype=object), term_ptr=term_ptr,
                 post_doc=post_doc, post_tf=post_tf, doc_len=doc_len, idf=idf, k1=np.float64(k1),
                 b=np.float64(b), avgdl=np.float64(avgdl), pickle_size=np.int64(st.st_size),
                 pickle_mtime_ns=np.int64(st.st_mtime_ns))
        os.replace(tmp, pickle_path + CSR_CACHE_SUFFIX)
    except OSError as e:   # read-only directory etc.: the cache is an optimisation only
        logger.info(f"CSR cache not written for {pickle_path}: {e}")


class Bm25Entry:
    def __init__(self, bm25, cache_for: Optional[str] = None):
        cached = _csr_cache_load(cache_for) if cache_for else None
        if cached is not None and float(cached["k1"]) == float(bm25.k1) \
                and float(cached["b"]) == float(bm25.b) and len(cached["doc_len"]) == len(bm25.doc_len):
            vocab = {tok: i for i, tok in enumerate(cached["vocab"].tolist())}
            self.index = engine.Bm25Index(cached["term_ptr"], cached["post_doc"], cached["post_tf"],
                                          cached["doc_len"], cached["idf"], float(cached["k1"]),
                                          float(cached["b"]), float(cached["avgdl"]), vocab=vocab)
        else:
            vocab, term_ptr, post_doc, post_tf, doc_len, idf = engine.invert_okapi(bm25)
            if cache_for:
                _csr_cache_save(cache_for, vocab, term_ptr, post_doc, post_tf, doc_len, idf,
                                bm25.k1, bm25.b, bm25.avgdl)
            self.index = engine.Bm25Index(term_ptr, post_doc, post_tf, doc_len, idf, bm25.k1,
                                          bm25.b, bm25.avgdl, vocab=vocab)
        self._masks: Dict[tuple, Tuple[object, object, int]] = {}
        self.fingerprint = bm25_fingerprint(bm25)

    def filter_mask(self, sections, filename_type_filter: str):
        """Mask over BM25 doc indices from each section's metadata["source"] (search_engine.py:224-231)."""
        # keyed by content, not by id(sections): an id can be reused by another list after GC
        n = len(sections)
        probe = tuple(sections[i].metadata.get("source", "") for i in sorted({0, n // 2, n - 1})) if n else ()
        key = (n, probe, filename_type_filter)
        hit = self._masks.get(key)
        if hit is None:
            sources = [s.metadata.get("source", "") for s in sections]
            mask = engine.prefix_mask(sources, filename_type_filter)
            if len(mask) != self.index.n_docs:
                raise ValueError("bm25_sections does not match the BM25 index")
            hit = (mask, device_words(engine.pack_mask(mask)), int(mask.sum()))
            self._masks[key] = hit
        return hit


_bm25: Dict[int, Tuple[weakref.ref, Bm25Entry]] = {}


def bm25_fingerprint(bm25) -> tuple:
    """What a cached device index was built from, as far as O(1) checks can tell: a BM25Okapi whose
    k1 / b / avgdl were changed, whose corpus grew, or whose idf table was replaced (the
    parameter sweep of src/processing/bm25_test.py does all of that on fresh objects) must not
    be searched through the old index.  In-place edits of single idf entries are not detectable."""
    idf = getattr(bm25, "idf", None)
    return (float(bm25.k1), float(bm25.b), float(bm25.avgdl), len(bm25.doc_len), id(idf),
            len(idf) if idf is not None else 0)


def resolve_bm25(bm25, cache_for: Optional[str] = None) -> Bm25Entry:
    entry = getattr(bm25, "_anr_entry", None)
    if isinstance(entry, Bm25Entry) and entry.fingerprint == bm25_fingerprint(bm25):
        return entry
    hit = _bm25.get(id(bm25))
    if hit is not None and hit[0]() is bm25 and hit[1].fingerprint == bm25_fingerprint(bm25):
        return hit[1]
    entry = Bm25Entry(bm25, cache_for=cache_for)
    try:
        bm25._anr_entry = entry
    except AttributeError:  # __slots__ / frozen object
        _bm25[id(bm25)] = (weakref.ref(bm25), entry)
        weakref.finalize(bm25, _bm25.pop, id(bm25), None)
    return entry
