"""Synthetic corpora in the reference's on-disk formats and in BASELINE.json's shapes.

Used by the tests, ``bench.py`` and ``__graft_entry__.smoke()``: unit-norm N(0,1) embedding
rows, Zipf(s) token documents and queries, guideline-style sources, plus writers for the real
formats the loaders read -- the SQLite ``chunks`` table (src/processing/create_database.py:59-64,
:113, with the ``url`` column added by notebooks/modify_db.ipynb) and the BM25 pickle
``{bm25, sections, section_ids}`` (src/processing/bm25_search.py:84-93).
No scoring happens here.
"""
from __future__ import annotations

import pickle
import sqlite3
import sys
import types
from typing import List, Optional, Sequence, Tuple

import numpy as np

SOURCE_PREFIXES = ("CG", "NG", "PH", "TA", "QS", "IPG")
SOURCE_PROBS = (0.34, 0.33, 0.09, 0.09, 0.08, 0.07)   # ~2/3 of rows pass the "CG,NG" filter


def unit_vectors(n: int, d: int, seed: int) -> np.ndarray:
    """[n, d] fp32, i.i.d. N(0,1) rows L2-normalised in fp32."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    return x


def zipf_probs(vocab: int, s: float) -> np.ndarray:
    p = 1.0 / np.power(np.arange(1, vocab + 1, dtype=np.float64), s)
    return p / p.sum()


def zipf_corpus(n_docs: int, vocab: int, s: float, seed: int, len_lo: int = 100,
                len_hi: int = 300) -> Tuple[np.ndarray, np.ndarray]:
    """(doc_ptr int64 [n_docs+1], tokens int32): doc i = tokens[doc_ptr[i]:doc_ptr[i+1]],
    token = Zipf rank - 1, document length ~ U[len_lo, len_hi)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(len_lo, len_hi, size=n_docs)
    doc_ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=doc_ptr[1:])
    cdf = np.cumsum(zipf_probs(vocab, s))
    tokens = np.searchsorted(cdf, rng.random(int(doc_ptr[-1])), side="right")
    return doc_ptr, np.minimum(tokens, vocab - 1).astype(np.int32)


def zipf_queries(n_queries: int, n_terms: int, vocab: int, s: float, seed: int,
                 skip_head: int = 0) -> np.ndarray:
    """[n_queries, n_terms] int32 term ranks - 1 from the same Zipf law (duplicates kept).
    skip_head > 0 resamples within ranks > skip_head (the "content-word" variant)."""
    rng = np.random.default_rng(seed)
    p = zipf_probs(vocab, s)
    if skip_head:
        p = p.copy()
        p[:skip_head] = 0.0
        p /= p.sum()
    cdf = np.cumsum(p)
    q = np.searchsorted(cdf, rng.random((n_queries, n_terms)), side="right")
    return np.minimum(q, vocab - 1).astype(np.int32)


def token_strings(ids: Sequence[int]) -> List[str]:
    return [f"t{int(i)}" for i in ids]


def doc_token_lists(doc_ptr: np.ndarray, tokens: np.ndarray) -> List[List[str]]:
    names = np.array([f"t{i}" for i in range(int(tokens.max()) + 1 if len(tokens) else 0)],
                     dtype=object)
    flat = names[tokens]
    return [list(flat[doc_ptr[i]:doc_ptr[i + 1]]) for i in range(len(doc_ptr) - 1)]


def sources(n: int, seed: int) -> List[str]:
    rng = np.random.default_rng(seed)
    kinds = rng.choice(len(SOURCE_PREFIXES), size=n, p=SOURCE_PROBS)
    nums = rng.integers(1, 250, size=n)
    return [f"{SOURCE_PREFIXES[k]}{m}" for k, m in zip(kinds, nums)]


def chunk_ids(n: int, srcs: Sequence[str]) -> List[str]:
    return [f"{srcs[i]}_Section {i}" for i in range(n)]


# ---------------------------------------------------------------------------------
# writers for the real on-disk formats
# ---------------------------------------------------------------------------------
def write_chunks_db(path: str, ids: Sequence[str], contents: Sequence[str], srcs: Sequence[str],
                    emb: np.ndarray, extra_rows: Sequence[tuple] = ()) -> None:
    """SQLite ``chunks(id, content, source, embedding BLOB fp32, url)``."""
    conn = sqlite3.connect(path)
    conn.execute("CREATE TABLE IF NOT EXISTS chunks (id TEXT PRIMARY KEY, content TEXT, "
                 "source TEXT, embedding BLOB)")
    conn.execute("ALTER TABLE chunks ADD COLUMN url TEXT")
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    conn.executemany(
        "INSERT INTO chunks (id, content, source, embedding, url) VALUES (?, ?, ?, ?, ?)",
        ((ids[i], contents[i], srcs[i], emb[i].tobytes(),
          "https://www.nice.org.uk/guidance/" + srcs[i].lower()) for i in range(len(ids))))
    if extra_rows:
        conn.executemany(
            "INSERT INTO chunks (id, content, source, embedding, url) VALUES (?, ?, ?, ?, ?)",
            extra_rows)
    conn.commit()
    conn.close()


class Document:
    """Pickles as ``langchain.schema.document.Document`` (what bm25_search.py:70 stores)."""

    def __init__(self, page_content: str, metadata: dict):
        self.page_content = page_content
        self.metadata = metadata


Document.__module__ = "langchain.schema.document"
Document.__qualname__ = "Document"


def _module_chain(name: str, created: list) -> types.ModuleType:
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        sub = ".".join(parts[:i])
        if sub not in sys.modules:
            sys.modules[sub] = types.ModuleType(sub)
            created.append(sub)
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[sub])
    return sys.modules[name]


def write_bm25_pickle(path: str, bm25, contents: Sequence[str], ids: Sequence[str],
                      srcs: Sequence[str], config: Optional[dict] = None) -> None:
    """``{bm25, sections, section_ids, config}`` with the class paths of a real index:
    the BM25 object is pickled as ``rank_bm25.BM25Okapi`` and the sections as langchain
    ``Document``s, whatever class actually built them here."""
    injected, created = [], []
    bm25_cls = type(bm25)
    saved = (bm25_cls.__module__, bm25_cls.__qualname__)
    try:
        for modname, cls, attr in (("rank_bm25", bm25_cls, "BM25Okapi"),
                                   ("langchain.schema.document", Document, "Document")):
            mod = _module_chain(modname, created)
            if getattr(mod, attr, None) is not cls:
                injected.append((mod, attr, getattr(mod, attr, None)))
                setattr(mod, attr, cls)
        bm25_cls.__module__, bm25_cls.__qualname__ = "rank_bm25", "BM25Okapi"
        sections = [Document(contents[i], {"id": ids[i], "source": srcs[i]})
                    for i in range(len(ids))]
        with open(path, "wb") as fh:
            pickle.dump({"bm25": bm25, "sections": sections, "section_ids": list(ids),
                         "config": config or {}}, fh)
    finally:
        bm25_cls.__module__, bm25_cls.__qualname__ = saved
        for mod, attr, old in injected:
            if old is None:
                delattr(mod, attr)
            else:
                setattr(mod, attr, old)
        for name in created:
            sys.modules.pop(name, None)


# ---------------------------------------------------------------------------------
# device-side generators for BASELINE-sized corpora (torch is the allocator / RNG here)
# ---------------------------------------------------------------------------------
def unit_vectors_torch(n: int, d: int, seed: int, device):
    """[n, d] fp32 unit rows generated on ``device`` in slabs (a 100M-row corpus cannot be
    built on the host)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    step = 1 << 18
    for r0 in range(0, n, step):
        slab = out[r0:r0 + step]
        slab.normal_(generator=g)
        slab /= slab.norm(dim=1, keepdim=True)
    return out


def zipf_postings_torch(n_docs: int, vocab: int, s: float, seed: int, device,
                        len_lo: int = 100, len_hi: int = 300):
    """Zipf corpus inverted on ``device``.  -> dict(term_ptr i64 [V+1], post_doc i32, post_tf
    i32, doc_len i32 [n_docs], nd i64 [V]); postings sorted by (term, doc).

    Built slab by slab (2^17 documents each: torch.unique handles < 2^31 keys): a slab's
    (term, doc, tf) triples come out sorted by (term, doc), slabs are in document order, so the
    postings of term t from slab j land at term_ptr[t] + (postings of t in slabs < j) + rank."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    lens = torch.randint(len_lo, len_hi, (n_docs,), generator=g, device=device)
    cdf = torch.cumsum(torch.from_numpy(zipf_probs(vocab, s)).to(device), 0)
    step = 1 << 17
    slabs = []                                   # (term i32, doc i32, tf i32, nd_slab i64)
    nd = torch.zeros(vocab, dtype=torch.int64, device=device)
    for d0 in range(0, n_docs, step):
        l = lens[d0:d0 + step]
        total = int(l.sum())
        u = torch.rand(total, generator=g, device=device, dtype=torch.float64)
        tok = torch.searchsorted(cdf, u, right=True).clamp_(max=vocab - 1)
        doc = torch.repeat_interleave(torch.arange(l.numel(), device=device), l)
        uniq, tf = torch.unique(tok * l.numel() + doc, return_counts=True)   # sorted (term, doc)
        del u, tok, doc
        term = torch.div(uniq, l.numel(), rounding_mode="floor")
        nd_slab = torch.bincount(term, minlength=vocab)
        slabs.append((term.to(torch.int32), (uniq - term * l.numel() + d0).to(torch.int32),
                      tf.to(torch.int32), nd_slab))
        nd += nd_slab
        del uniq, term
    term_ptr = torch.zeros(vocab + 1, dtype=torch.int64, device=device)
    term_ptr[1:] = torch.cumsum(nd, 0)
    nnz = int(term_ptr[-1])
    post_doc = torch.empty(nnz, dtype=torch.int32, device=device)
    post_tf = torch.empty(nnz, dtype=torch.int32, device=device)
    base = term_ptr[:-1].clone()                 # next free slot of every term
    for term, doc, tf, nd_slab in slabs:
        start = torch.cumsum(nd_slab, 0) - nd_slab           # first index of each term in the slab
        t = term.long()
        dest = base[t] + (torch.arange(t.numel(), device=device) - start[t])
        post_doc[dest] = doc
        post_tf[dest] = tf
        base += nd_slab
        del dest, t
    del slabs
    return dict(term_ptr=term_ptr, post_doc=post_doc, post_tf=post_tf,
                doc_len=lens.to(torch.int32), nd=nd)


def idf_from_counts(n_docs_total: int, nd: np.ndarray, epsilon: float) -> np.ndarray:
    """float64 idf per term id from document frequencies, rank-bm25 0.2.2 rule: raw
    ln(N - nd + .5) - ln(nd + .5); terms with raw < 0 get epsilon * mean(raw over present
    terms); absent terms (nd == 0) are not in the vocabulary -> idf 0."""
    nd = np.asarray(nd, dtype=np.float64)
    present = nd > 0
    raw = np.zeros_like(nd)
    raw[present] = np.log(n_docs_total - nd[present] + 0.5) - np.log(nd[present] + 0.5)
    floor = epsilon * (raw[present].sum() / max(int(present.sum()), 1))
    return np.where(present & (raw < 0), floor, raw)
