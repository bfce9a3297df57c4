"""CPU: host-side logic and the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import importlib
import os
import re
import sys

import numpy as np
import pandas as pd
import pytest

from helpers import ROOT, csr_from_case, synth
from oracle import csr, retrieval

engine = importlib.import_module("a-nice-rag_b200.engine")
native = importlib.import_module("a-nice-rag_b200.native")
registry = importlib.import_module("a-nice-rag_b200.registry")


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "anr_b200.h")).read()
    declared = set(re.findall(r"\b(anr_[a-z0-9_]+)\s*\(", header))
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    lib = native.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.anr_abi_version() == 1


def test_set_option_knows_its_keys():
    """anr_set_option is host-only state (no device call): the launch mask is accepted, an unknown
    key is ANR_ERR_INVALID with a message, NULL is rejected."""
    native.call("anr_set_option", b"pdl", 0)
    native.call("anr_set_option", b"pdl", 7)
    with pytest.raises(native.AnrError) as err:
        native.call("anr_set_option", b"no_such_knob", 1)
    assert "unknown key" in str(err.value)
    with pytest.raises(native.AnrError):
        native.call("anr_set_option", None, 1)


def test_no_device_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    handle = ctypes.c_void_p()
    rc = native.load().anr_ctx_create(0, ctypes.byref(handle))
    assert rc == 3 and b"no CPU implementation" in native.load().anr_last_error()
    with pytest.raises(native.AnrError):
        engine.DenseIndex(np.zeros((4, 8), dtype=np.float32))


def test_pack_mask_layout():
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 1000):
        mask = rng.random(n) < 0.4
        words = engine.pack_mask(mask)
        assert words.dtype == np.uint32 and len(words) == (n + 31) // 32
        for i in range(n):
            assert bool((words[i >> 5] >> (i & 31)) & 1) == bool(mask[i])


def test_prefix_mask_matches_reference_semantics():
    srcs = ["CG12", "cg7", "NG1", "PH2", None, "", "XCG1", " ng3", float("nan")]
    for flt in ("CG, ng", "cg", "ZZ", "C"):
        assert engine.prefix_mask(srcs, flt).tolist() == retrieval.filter_mask(srcs, flt).tolist()


def test_invert_okapi_equals_oracle_inversion(small_case):
    ix, okapi = csr_from_case(small_case)
    vocab, term_ptr, post_doc, post_tf, doc_len, idf = engine.invert_okapi(okapi)
    want = csr.from_okapi(okapi)
    assert vocab == want.vocab
    assert np.array_equal(term_ptr, want.term_ptr) and np.array_equal(post_doc, want.post_doc)
    assert np.array_equal(post_tf, want.post_tf) and np.array_equal(idf, want.idf)
    assert np.array_equal(doc_len, want.doc_len)
    for t in range(len(term_ptr) - 1):          # ascending documents inside every term
        seg = post_doc[term_ptr[t]:term_ptr[t + 1]]
        assert np.all(np.diff(seg) > 0)


def test_pack_queries_csr():
    terms, offsets = engine.Bm25Index.pack_queries([[3, 3, -1], [], [7]])
    assert terms.tolist() == [3, 3, -1, 7] and offsets.tolist() == [0, 3, 3, 4]


def test_database_manager_contract_without_gpu(tmp_path, caplog):
    pkg = importlib.import_module("a-nice-rag_b200")
    n, d = 50, 12
    emb = synth.unit_vectors(n, d, seed=1)
    srcs = synth.sources(n, seed=2)
    ids = synth.chunk_ids(n, srcs)
    db = str(tmp_path / "c.db")
    bad = [("bad1", "x", "CG1", b"\x00\x01\x02", "u"), ("bad2", "x", "CG1", None, "u")]
    synth.write_chunks_db(db, ids, [f"doc {i}" for i in range(n)], srcs, emb, extra_rows=bad)
    dm = pkg.DatabaseManager()
    df = dm.load_embeddings_from_sql(db, "m")
    assert list(df.columns) == ["id", "document", "source", "embedding", "url"]
    assert len(df) == n and list(df["id"]) == ids            # invalid rows skipped
    assert all(isinstance(e, np.ndarray) and e.dtype == np.float32 for e in df["embedding"])
    assert np.array_equal(np.stack(df["embedding"].values), emb)
    assert df["url"][0] == "https://www.nice.org.uk/guidance/" + srcs[0].lower()
    assert dm.load_embeddings_from_sql(db, "m") is df        # cached
    with pytest.raises(FileNotFoundError):
        dm.load_embeddings_from_sql(str(tmp_path / "missing.db"))
    with pytest.raises(FileNotFoundError):
        dm.load_bm25_from_pickle(str(tmp_path / "missing.pkl"))
    # derived frames are recognised as row subsets of the registered one
    entry, subset = registry.resolve_frame(df)
    assert subset is None and entry.n == n
    sub = df[df["source"].str.startswith("CG")].copy()
    e2, rows = registry.resolve_frame(sub)
    assert e2 is entry and rows.tolist() == sub.index.tolist()
    empty_db = str(tmp_path / "e.db")
    synth.write_chunks_db(empty_db, [], [], [], np.zeros((0, 4), dtype=np.float32))
    assert dm.load_embeddings_from_sql(empty_db).empty
    # Same ids, OTHER vectors: the reference re-stacks df["embedding"] on every call
    # (search_engine.py:80), so neither a copy with a new embedding column nor the loaded frame
    # with its column replaced may resolve to the matrix that was uploaded at load time.
    renorm = df.copy()
    renorm["embedding"] = [2.0 * e for e in df["embedding"]]
    e3, rows3 = registry.resolve_frame(renorm)
    assert e3 is not entry and rows3 is None
    assert np.array_equal(e3.packed, 2.0 * emb)
    assert registry.resolve_frame(renorm)[0] is e3            # remembered by identity
    renorm["embedding"] = [3.0 * e for e in df["embedding"]]   # ...until its vectors change again
    e4, _ = registry.resolve_frame(renorm)
    assert e4 is not e3 and np.array_equal(e4.packed, 3.0 * emb)
    sub2 = df[df["source"].str.startswith("CG")].copy()
    sub2["embedding"] = [-e for e in sub2["embedding"]]
    e5, rows5 = registry.resolve_frame(sub2)
    assert e5 is not entry and rows5 is None and np.array_equal(e5.packed, -emb[sub2.index])
    assert registry.resolve_frame(df)[0] is entry               # the loaded frame is unaffected


def test_bm25_entry_follows_parameter_changes(small_case, monkeypatch):
    """A BM25Okapi whose k1 / b / idf table changed after its device index was built must get a
    new index (the k1 / b / epsilon sweep of src/processing/bm25_test.py:180-315)."""
    _, okapi = csr_from_case(small_case)
    built = []

    class FakeIndex:
        def __init__(self, *a, **kw):
            built.append(a[5:8])          # (k1, b, avgdl)
            self.n_docs = len(a[3])
    monkeypatch.setattr(engine, "Bm25Index", FakeIndex)
    e1 = registry.resolve_bm25(okapi)
    assert registry.resolve_bm25(okapi) is e1 and len(built) == 1
    okapi.k1 = 1.2
    e2 = registry.resolve_bm25(okapi)
    assert e2 is not e1 and built[-1][0] == 1.2
    okapi.idf = dict(okapi.idf)           # a recomputed idf table (new epsilon floor)
    assert registry.resolve_bm25(okapi) is not e2 and len(built) == 3
    # the filter-mask cache is keyed by content, not by id(sections)
    srcs = list(small_case["sources"])
    sections = [synth.Document("", {"id": str(i), "source": s}) for i, s in enumerate(srcs)]
    monkeypatch.setattr(registry, "device_words", lambda w: None)
    m1 = e2.filter_mask(sections, "CG")
    assert e2.filter_mask(list(sections), "CG") is m1
    other = [synth.Document("", {"id": str(i), "source": "NG1"}) for i in range(len(srcs))]
    assert e2.filter_mask(other, "CG") is not m1 and e2.filter_mask(other, "CG")[2] == 0


def test_bm25_pickle_roundtrip_with_shims(tmp_path, small_case):
    pkg = importlib.import_module("a-nice-rag_b200")
    _, okapi = csr_from_case(small_case)
    n = len(okapi.doc_freqs)
    srcs = list(small_case["sources"])
    ids = synth.chunk_ids(n, srcs)
    pkl = str(tmp_path / "b.pkl")
    synth.write_bm25_pickle(pkl, okapi, [""] * n, ids, srcs)
    assert "rank_bm25" not in sys.modules or hasattr(sys.modules["rank_bm25"], "__file__")
    bm25, sections, section_ids = pkg.DatabaseManager().load_bm25_from_pickle(pkl)
    assert type(bm25).__name__ == "BM25Okapi" and type(bm25).__module__ == "rank_bm25"
    assert bm25.idf == okapi.idf and bm25.doc_len == okapi.doc_len and bm25.avgdl == okapi.avgdl
    assert section_ids == ids and sections[3].metadata == {"id": ids[3], "source": srcs[3]}


def test_reference_named_modules_resolve_to_the_product():
    src = os.path.join(ROOT, "a-nice-rag_b200", "src")
    saved = {k: sys.modules.pop(k, None) for k in ("search_engine", "database_manager", "config",
                                                   "processing", "processing.preprocess_bm25")}
    sys.path.insert(0, src)
    try:
        se = importlib.import_module("search_engine")
        dm = importlib.import_module("database_manager")
        cfg = importlib.import_module("config")
        pp = importlib.import_module("processing.preprocess_bm25")
        pkg = importlib.import_module("a-nice-rag_b200")
        assert se.SearchEngine is pkg.SearchEngine and dm.DatabaseManager is pkg.DatabaseManager
        assert cfg.Config.DEFAULT_MODEL_WEIGHTS == {"voyage-3-large": 5.0, "BM25": 1.0,
                                                    "text-embedding-3-large": 0.0,
                                                    "voyage-3.5": 0.0, "Qwen3": 0.0}
        assert pp.preprocess_text("The 12 patients with Hypertension!") == ["patients", "hypertension"]
        with pytest.raises(ValueError):
            cfg.Config.get_source_config("nope")
        assert cfg.Config.get_source_config("NICE").voyage_db_path.endswith("2048.db")
    finally:
        sys.path.remove(src)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_signatures_match_the_reference():
    import inspect
    pkg = importlib.import_module("a-nice-rag_b200")
    se = pkg.SearchEngine
    want = {
        "similarity_search_with_embedding": ["self", "query_embedding", "df", "model_name",
                                             "similarity_k", "filename_type_filter"],
        "similarity_search": ["self", "query_text", "df", "model_name", "similarity_k",
                              "filename_type_filter", "query_embedding"],
        "bm25_search": ["self", "query_text", "bm25", "bm25_sections", "bm25_section_ids",
                        "similarity_k", "filename_type_filter", "use_lemmatized"],
        "bm25_search_preprocessed": ["self", "query_tokens", "bm25", "bm25_sections",
                                     "bm25_section_ids", "similarity_k", "filename_type_filter"],
        "weighted_reciprocal_rank_fusion": ["self", "ranked_lists", "model_weights", "k"],
        "rerank_documents": ["self", "query_text", "documents", "reranker_model", "reranker_top_k"],
    }
    for name, params in want.items():
        sig = inspect.signature(getattr(se, name))
        assert list(sig.parameters) == params, name
    assert inspect.signature(se.weighted_reciprocal_rank_fusion).parameters["k"].default == 50
    assert inspect.signature(se.similarity_search).parameters["similarity_k"].default == 25
    assert se(None, None).bm25_search_preprocessed([], None, None, None) == []
    assert se(None, None).weighted_reciprocal_rank_fusion([], {}) == []


def test_shard_ranges_cover_without_overlap():
    sharded = importlib.import_module("a-nice-rag_b200.sharded")
    for n in (0, 1, 7, 1_000_000, 100_000_003):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_idf_from_counts_matches_okapi(small_case):
    ix, okapi = csr_from_case(small_case)
    nd = np.diff(ix.term_ptr)
    idf = synth.idf_from_counts(len(okapi.doc_freqs), nd, 0.05)
    np.testing.assert_allclose(idf, ix.idf, rtol=1e-12, atol=1e-15)


@pytest.mark.skipif(not os.path.exists("/root/reference/data/test_queries_bm25.csv"),
                    reason="reference tokeniser goldens not mounted")
def test_offline_tokeniser_reproduces_the_reference_goldens():
    """The only input->output goldens the reference ships: 9 609 + 8 168
    (query, tokens_regular) rows written by its NLTK pipeline (preprocess_bm25.py:33-52).  The
    offline tokeniser must reproduce every one of them (lemmatised tokens need WordNet)."""
    import ast
    pp = importlib.import_module("a-nice-rag_b200.processing.preprocess_bm25")
    for name in ("suggested_queries_bm25_preprocessed.csv", "test_queries_bm25.csv"):
        df = pd.read_csv(os.path.join("/root/reference/data", name))
        bad = [q for q, t in zip(df["query"], df["tokens_regular"])
               if pp.preprocess_text(q) != ast.literal_eval(t)]
        assert not bad, (name, len(bad), bad[:3])


def test_offline_tokeniser_on_frozen_golden_sample():
    import ast
    import json
    pp = importlib.import_module("a-nice-rag_b200.processing.preprocess_bm25")
    with open(os.path.join(ROOT, "tests", "golden", "tokeniser_sample.json")) as fh:
        rows = json.load(fh)["rows"]
    assert len(rows) == 60
    for row in rows:
        assert pp.preprocess_text(row["query"]) == ast.literal_eval(row["tokens_regular"]), row
        assert (pp.preprocess_text(row["query"], use_lemmatization=True)
                == ast.literal_eval(row["tokens_lemmatized"])), row


def _lemma_fixture_rows(name):
    import ast
    df = pd.read_csv(os.path.join("/root/reference/data", name))
    return [(q, ast.literal_eval(a), ast.literal_eval(b))
            for q, a, b in zip(df["query"], df["tokens_regular"], df["tokens_lemmatized"])]


@pytest.mark.skipif(not os.path.exists("/root/reference/data/test_queries_bm25.csv"),
                    reason="reference tokeniser goldens not mounted")
def test_offline_lemmatiser_reproduces_the_reference_goldens():
    """``tokens_lemmatized`` of all 17 777 fixture rows (the reference's BM25 default,
    search_engine.py:253,258 -> preprocess_bm25.py:48-50) through the offline lemma path; and the
    committed table is exactly what oracle/make_lemma_table.py distils from those files today."""
    pp = importlib.import_module("a-nice-rag_b200.processing.preprocess_bm25")
    for name in ("suggested_queries_bm25_preprocessed.csv", "test_queries_bm25.csv"):
        rows = _lemma_fixture_rows(name)
        bad = [q for q, _, lem in rows if pp.preprocess_text(q, use_lemmatization=True) != lem]
        assert not bad, (name, len(bad), bad[:3])
    import json
    from importlib import util as import_util
    spec = import_util.spec_from_file_location(
        "_mk_lemma", os.path.join(ROOT, "oracle", "make_lemma_table.py"))
    mk = import_util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    with open(mk.OUT) as fh:
        assert json.load(fh) == mk.build()


@pytest.mark.skipif(not os.path.exists("/root/reference/data/test_queries_bm25.csv"),
                    reason="reference tokeniser goldens not mounted")
def test_offline_lemmatiser_generalises_to_unseen_tokens():
    """Held-out rate of the rule path: the table distilled from ONE fixture file only, applied
    to the other file (320 tokens it has never seen).  The bar written in DESIGN.md:
    >= 99.9 % of rows, >= 97 % of unseen tokens."""
    pp = importlib.import_module("a-nice-rag_b200.processing.preprocess_bm25")
    train = _lemma_fixture_rows("test_queries_bm25.csv")
    test = _lemma_fixture_rows("suggested_queries_bm25_preprocessed.csv")
    changed, keep = {}, set()
    for _, reg, lem in train:
        for a, b in zip(reg, lem):
            if a != b:
                changed[a] = b
            else:
                keep.add(a)
    keep = frozenset(keep)
    rows_ok = unseen = unseen_ok = 0
    for _, reg, lem in test:
        got = [pp.offline_lemma(t, changed, keep) for t in reg]
        rows_ok += got == lem
        for t, g, want in zip(reg, got, lem):
            if t not in changed and t not in keep:
                unseen += 1
                unseen_ok += g == want
    assert unseen >= 300
    assert rows_ok / len(test) >= 0.999, rows_ok / len(test)
    assert unseen_ok / unseen >= 0.97, unseen_ok / unseen


def test_csr_cache_roundtrip_and_invalidation(tmp_path, small_case):
    _, okapi = csr_from_case(small_case)
    n = len(okapi.doc_freqs)
    srcs = list(small_case["sources"])
    pkl = str(tmp_path / "b.pkl")
    synth.write_bm25_pickle(pkl, okapi, [""] * n, synth.chunk_ids(n, srcs), srcs)
    assert registry._csr_cache_load(pkl) is None
    vocab, term_ptr, post_doc, post_tf, doc_len, idf = engine.invert_okapi(okapi)
    registry._csr_cache_save(pkl, vocab, term_ptr, post_doc, post_tf, doc_len, idf, okapi.k1,
                             okapi.b, okapi.avgdl)
    got = registry._csr_cache_load(pkl)
    assert got is not None and got["vocab"].tolist() == list(vocab.keys())
    for name, want in (("term_ptr", term_ptr), ("post_doc", post_doc), ("post_tf", post_tf),
                       ("doc_len", doc_len), ("idf", idf)):
        assert np.array_equal(got[name], want), name
    assert float(got["avgdl"]) == okapi.avgdl
    os.utime(pkl, ns=(1, 1))                      # the pickle changed: the cache is stale
    assert registry._csr_cache_load(pkl) is None


def test_loader_streams_into_one_matrix_and_handles_ragged_tables(tmp_path):
    pkg = importlib.import_module("a-nice-rag_b200")
    n, d = 40_000, 16                              # several fetchmany() batches
    emb = synth.unit_vectors(n, d, seed=3)
    srcs = synth.sources(n, seed=4)
    ids = synth.chunk_ids(n, srcs)
    db = str(tmp_path / "big.db")
    synth.write_chunks_db(db, ids, [""] * n, srcs, emb)
    df = pkg.DatabaseManager().load_embeddings_from_sql(db)
    entry, subset = registry.resolve_frame(df)
    assert subset is None and entry.packed.shape == (n, d) and np.array_equal(entry.packed, emb)
    assert all(e.base is not None for e in df["embedding"].iloc[:5])      # views, not copies
    # a table whose rows disagree on the width keeps the reference's frame (np.stack fails later)
    db2 = str(tmp_path / "ragged.db")
    odd = [("odd", "x", "CG1", np.zeros(d + 4, dtype=np.float32).tobytes(), "u")]
    synth.write_chunks_db(db2, ids[:10], [""] * 10, srcs[:10], emb[:10], extra_rows=odd)
    df2 = pkg.DatabaseManager().load_embeddings_from_sql(db2)
    assert len(df2) == 11 and df2["embedding"].iloc[10].shape == (d + 4,)
    assert registry.ATTR_KEY not in df2.attrs
    assert np.array_equal(np.stack(df2["embedding"].iloc[:10].values), emb[:10])


def test_bulk_blob_loader_abi_and_paths(tmp_path, monkeypatch):
    """anr_sqlite_read_blobs (host-only entry point) and the two loader paths:
    a uniform table goes through it, odd tables fall back to the row-by-row loop, and both give
    the frame the UNMODIFIED reference loader gives where it is mounted."""
    import ctypes as C
    pkg = importlib.import_module("a-nice-rag_b200")
    native = pkg.native
    n, d = 3000, 20
    emb = synth.unit_vectors(n, d, seed=5)
    srcs = synth.sources(n, seed=6)
    ids = synth.chunk_ids(n, srcs)
    db = str(tmp_path / "u.db")
    synth.write_chunks_db(db, ids, [f"doc {i}" for i in range(n)], srcs, emb)

    def read(sql, row_bytes, max_rows):
        dst = np.zeros((max(max_rows, 1), row_bytes // 4), dtype=np.float32)
        rowids = np.zeros(max(max_rows, 1), dtype=np.int64)
        got, uniform = C.c_int64(), C.c_int32()
        native.call("anr_sqlite_read_blobs", db.encode(), sql, dst.ctypes.data, row_bytes, max_rows,
                    rowids.ctypes.data, C.byref(got), C.byref(uniform))
        return dst, rowids, got.value, uniform.value

    dst, rowids, got, uniform = read(b"SELECT rowid, embedding FROM chunks", d * 4, n)
    assert (got, uniform) == (n, 1) and np.array_equal(dst, emb)
    assert rowids.tolist() == list(range(1, n + 1))
    assert read(b"SELECT rowid, embedding FROM chunks", d * 4, n - 1)[2:] == (n - 1, 0)   # table too long
    assert read(b"SELECT rowid, embedding FROM chunks", d * 4 + 4, n)[2:] == (0, 0)        # other width
    assert read(b"SELECT rowid, id FROM chunks", d * 4, n)[2:] == (0, 0)                   # not a BLOB
    with pytest.raises(native.AnrError, match="sqlite3_prepare_v2"):
        read(b"SELECT rowid, nothing FROM chunks", d * 4, n)
    with pytest.raises(native.AnrError, match="rowid, blob"):
        read(b"SELECT embedding FROM chunks", d * 4, n)

    # the loader takes the bulk path for this table ...
    calls = []
    real = pkg.DatabaseManager._load_row_by_row
    monkeypatch.setattr(pkg.DatabaseManager, "_load_row_by_row",
                        lambda self, *a: calls.append(1) or real(self, *a))
    df = pkg.DatabaseManager().load_embeddings_from_sql(db, "m")
    assert not calls and list(df["id"]) == ids and np.array_equal(np.stack(df["embedding"].values), emb)
    assert list(df["document"]) == [f"doc {i}" for i in range(n)] and list(df["source"]) == srcs
    # ... and the row-by-row path when a row is odd (NULL embedding in the middle of the table)
    db2 = str(tmp_path / "odd.db")
    synth.write_chunks_db(db2, ids, [""] * n, srcs, emb, extra_rows=[("nul", "x", "CG1", None, "u")])
    df2 = pkg.DatabaseManager().load_embeddings_from_sql(db2, "m")
    assert calls and list(df2["id"]) == ids
    if reference_available():
        from oracle import reference_loader
        ref = reference_loader.load_reference()
        for path, ours in ((db, df), (db2, df2)):
            rdf = ref.DatabaseManager().load_embeddings_from_sql(path, "m")
            assert list(rdf.columns) == list(ours.columns) and len(rdf) == len(ours)
            for col in ("id", "document", "source", "url"):
                assert rdf[col].tolist() == ours[col].tolist()
            assert all(np.array_equal(a, b) for a, b in zip(rdf["embedding"], ours["embedding"]))


def test_config_values_equal_the_reference_module(pkg):
    """Every public value of the reference's src/config.py (weights, every SourceConfig field,
    enum members, the ValueError text) against the product's config."""
    if not reference_available():
        pytest.skip("reference sources not mounted")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_config", "/root/reference/src/config.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ours = pkg.config
    assert ours.Config.DEFAULT_MODEL_WEIGHTS == ref.Config.DEFAULT_MODEL_WEIGHTS
    assert list(ours.Config.DEFAULT_MODEL_WEIGHTS) == list(ref.Config.DEFAULT_MODEL_WEIGHTS)
    assert [(m.name, m.value) for m in ours.InfoSource] == [(m.name, m.value) for m in ref.InfoSource]
    fields = ("db_path", "bm25_path", "context_description", "not_found_message", "voyage_db_path",
              "voyage_3_5_db_path", "openai_db_path", "qwen_db_path")
    for member in ref.InfoSource:
        theirs = ref.Config.get_source_config(member.value.upper())
        mine = ours.Config.get_source_config(member.value.upper())
        assert [getattr(mine, f) for f in fields] == [getattr(theirs, f) for f in fields]
    assert ours.SourceConfig("a.db", "b.pkl", "c", "d").voyage_db_path == \
        ref.SourceConfig("a.db", "b.pkl", "c", "d").voyage_db_path == "a.db"
    for bad in ("nope", "NICE "):
        with pytest.raises(ValueError) as e_ref:
            ref.Config.get_source_config(bad)
        with pytest.raises(ValueError) as e_ours:
            ours.Config.get_source_config(bad)
        assert str(e_ours.value) == str(e_ref.value)


def reference_available() -> bool:
    from oracle import reference_loader
    return reference_loader.available()


# ---------------------------------------------------------------------------------------
# batched orchestrator: input validation and id space (no GPU needed)
def test_retrieve_documents_batch_validates_like_the_reference(pkg):
    import types
    batch = importlib.import_module("a-nice-rag_b200.batch_retrieval")
    system = types.SimpleNamespace(config=pkg.Config(), embeddings_data={}, bm25_data={},
                                   search_engine=None)
    q = np.zeros((2, 8), dtype=np.float32)
    with pytest.raises(ValueError, match="cannot be empty"):
        batch.retrieve_documents_batch(system, {})
    with pytest.raises(ValueError, match="must be a numpy array"):
        batch.retrieve_documents_batch(system, {"voyage-3-large": [[0.0]]})
    with pytest.raises(ValueError, match="cannot be empty"):
        batch.retrieve_documents_batch(system, {"voyage-3-large": np.zeros((0, 8), np.float32)})
    with pytest.raises(ValueError, match="positive integers"):
        batch.retrieve_documents_batch(system, {"voyage-3-large": q}, similarity_k=0)
    with pytest.raises(ValueError, match="Invalid info_source"):
        batch.retrieve_documents_batch(system, {"voyage-3-large": q}, info_source="nowhere")
    with pytest.raises(ValueError, match="one embedding per query"):
        batch.retrieve_documents_batch(system, {"voyage-3-large": q, "Qwen3": np.zeros((3, 8), np.float32)})
    # no data for the source: one empty result per query (query_rag_retrieval.py:183-185)
    assert batch.retrieve_documents_batch(system, {"voyage-3-large": q}) == [[], []]


def test_batch_id_space_is_first_seen_first_and_shared():
    batch = importlib.import_module("a-nice-rag_b200.batch_retrieval")
    space = batch._IdSpace()
    a = space.codes(["x", "y", "x", "z"])
    b = space.codes(["z", "w", "y"])
    assert a.tolist() == [0, 1, 0, 2] and b.tolist() == [2, 3, 1]
    assert space.names == ["x", "y", "z", "w"]


def test_reference_arm_contract_under_torchrun(tmp_path):
    """`bench.py --impl reference` launched like the driver launches it for N > 1: rank 0 alone
    runs the CPU port and prints ONE JSON line carrying the reference-arm keys; the other rank
    exits 0 without work.  (Tiny corpus: this checks the contract, not the number.)"""
    import json
    import subprocess
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
           "--impl", "reference", "--gpus", "2", "--chunks", "4000", "--steps", "2", "--warmup", "1"]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["n_gpus"] == 2 and rec["steps"] == 2
    assert rec["metric"].startswith("hybrid queries/sec") and rec["unit"] == "queries/s"
    assert rec["higher_is_better"] is True and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"] == {"value": rec["value"], "unit": "queries/s", "h2d_bytes_per_step": 0,
                          "d2h_bytes_per_step": 0}
    assert rec["config"]["batch"] == 64 and "4000 chunks" in rec["config"]["workload"]
    # the arm pins the BLAS pool itself (torchrun exports OMP_NUM_THREADS=1), emits the config
    # object of the CUDA arm key for key, and never maps the CUDA library
    assert rec["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    sys.path.insert(0, ROOT)
    import bench
    ns = bench.argparse.Namespace(chunks=4000, vocab=50_000, batch=64, gpus=2)
    assert rec["config"] == bench.headline_config(ns)
    probe = subprocess.run(
        [sys.executable, "-c",
         "import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--chunks', '2000', "
         "'--steps', '1', '--warmup', '0']; runpy.run_path(r'%s', run_name='__main__'); "
         "print('LIBS', [l.split()[-1] for l in open('/proc/self/maps') if 'libanr_b200' in l])"
         % os.path.join(ROOT, "bench.py")], capture_output=True, text=True, timeout=600,
        cwd=str(tmp_path))
    assert probe.returncode == 0, probe.stderr[-2000:]
    assert "LIBS []" in probe.stdout, probe.stdout[-500:]
