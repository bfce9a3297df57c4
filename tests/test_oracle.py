"""CPU: the oracle restatements against the golden vectors produced by the reference's own
code (oracle/make_golden.py), against each other, and -- where /root/reference exists --
against the reference imported verbatim."""
import numpy as np
import pytest

from helpers import (WEIGHTS, WRRF_K, check_ids_only, check_topk, csr_from_case, filter_mask,
                     okapi_from_case, synth, tag)
from oracle import bm25_okapi, csr, pipeline, reference_loader, retrieval

FILTERS = (None, "CG,NG", "cg", "ZZ")
KS = (10, 100, 3000)


@pytest.fixture(scope="module")
def built(small_case):
    ix, okapi = csr_from_case(small_case)
    return ix, okapi


def test_golden_loader_facts(small_case, built):
    _, okapi = built
    assert list(small_case["idf_vocab"]) == list(okapi.idf.keys())
    assert np.array_equal(small_case["idf_values"], np.array(list(okapi.idf.values())))
    assert float(small_case["avgdl"]) == okapi.avgdl
    # the case plants what it promises: negative raw idf (floored to epsilon * average_idf)
    n = len(okapi.doc_freqs)
    nd0 = sum(1 for d in okapi.doc_freqs if "t0" in d)
    assert nd0 > n / 2 and okapi.idf["t0"] == okapi.epsilon * okapi.average_idf


def test_csr_equals_literal_okapi_bitwise(small_case, built):
    ix, okapi = built
    for q in range(small_case["term_queries"].shape[0]):
        terms = small_case["term_queries"][q]
        lit = okapi.get_scores(synth.token_strings(terms))
        ids = [int(t) if t < ix.idf.shape[0] else -1 for t in terms]
        assert np.array_equal(csr.scores(ix, ids), lit), q
        assert np.array_equal(lit, small_case["bm25_all_scores"][q]), q


def test_from_okapi_inversion_matches(built):
    ix, okapi = built
    inv = csr.from_okapi(okapi)
    q = ["t0", "t3", "t3", "t17", "nope"]
    assert np.array_equal(csr.scores_for_tokens(inv, q), okapi.get_scores(q))


@pytest.mark.parametrize("flt", FILTERS)
@pytest.mark.parametrize("k", KS)
def test_dense_restatement_vs_golden(small_case, flt, k):
    emb, queries = small_case["emb"], small_case["queries"]
    mask = None if flt is None else filter_mask(small_case, flt)
    for q in range(queries.shape[0]):
        want_n = int(small_case[f"dense_counts_{tag(flt)}_k{k}"][q])
        want_ids = small_case[f"dense_ids_{tag(flt)}_k{k}"][q, :want_n]
        want_sc = small_case[f"dense_scores_{tag(flt)}_k{k}"][q, :want_n]
        rows, scores = retrieval.dense_topk(queries[q], emb, k, mask)
        full = retrieval.dense_scores(queries[q], emb)
        check_topk(rows, scores, want_ids, want_sc, full, f"dense q{q} {flt} k{k}")


@pytest.mark.parametrize("flt", FILTERS)
@pytest.mark.parametrize("k", KS)
def test_bm25_restatement_vs_golden(small_case, built, flt, k):
    ix, _ = built
    srcs = list(small_case["sources"])
    for q in range(small_case["term_queries"].shape[0]):
        terms = [int(t) if t < ix.idf.shape[0] else -1 for t in small_case["term_queries"][q]]
        all_scores = csr.scores(ix, terms)
        got = retrieval.bm25_topk(all_scores, k, srcs, flt)
        want_n = int(small_case[f"bm25_counts_{tag(flt)}_k{k}"][q])
        want = small_case[f"bm25_ids_{tag(flt)}_k{k}"][q, :want_n]
        if flt:   # stable sort: identical, ties included
            assert np.array_equal(got, want), (q, flt, k)
        else:
            check_ids_only(got, want, all_scores, f"bm25 q{q} k{k}")


@pytest.mark.parametrize("flt", (None, "CG,NG"))
def test_wrrf_restatement_vs_golden_bitwise(small_case, flt):
    k = 10
    for q in range(small_case["queries"].shape[0]):
        dn = int(small_case[f"dense_counts_{tag(flt)}_k{k}"][q])
        bn = int(small_case[f"bm25_counts_{tag(flt)}_k{k}"][q])
        d = small_case[f"dense_ids_{tag(flt)}_k{k}"][q, :dn].tolist()
        b = small_case[f"bm25_ids_{tag(flt)}_k{k}"][q, :bn].tolist()
        fused = retrieval.weighted_rrf([(d, "voyage-3-large"), (b, "BM25")], WEIGHTS, WRRF_K)
        fn = int(small_case[f"fused_counts_{tag(flt)}_k{k}"][q])
        assert [i for i, _ in fused] == small_case[f"fused_ids_{tag(flt)}_k{k}"][q, :fn].tolist()
        assert [s for _, s in fused] == small_case[f"fused_scores_{tag(flt)}_k{k}"][q, :fn].tolist()


def test_pipeline_matches_golden_fusion(small_case, built):
    ix, _ = built
    k = 10
    for q in range(small_case["queries"].shape[0]):
        terms = [int(t) if t < ix.idf.shape[0] else -1 for t in small_case["term_queries"][q]]
        res = pipeline.hybrid_query(small_case["queries"][q], small_case["emb"], ix, terms, k, k,
                                    WEIGHTS, WRRF_K, 2 * k)
        # fused list is exact GIVEN the lists; lists themselves are checked above
        pipeline.check_fused([i for i, _ in res["fused"]], [s for _, s in res["fused"]],
                             res["dense_ids"], res["bm25_ids"], (5.0, 1.0), WRRF_K, 2 * k)


def test_bm25_okapi_published_known_answer():
    """The one known-answer vector published for the third-party arithmetic: the usage example in
    the README of rank_bm25 (dorianbrown/rank_bm25, 0.2.x) -- default k1=1.5, b=0.75,
    epsilon=0.25 -- prints ``array([0., 0.93729472, 0.])`` for the query "windy London" and
    returns the London sentence from ``get_top_n``.  (The package itself is not available
    offline; the vector is quoted from its documentation, 8 significant digits.)  "is" occurs in
    two of the three documents, so the same corpus also exercises the negative-idf floor."""
    corpus = ["Hello there good man!", "It is quite windy in London", "How is the weather today?"]
    tokenized = [doc.split(" ") for doc in corpus]
    bm25 = bm25_okapi.BM25Okapi(tokenized)
    scores = bm25.get_scores("windy London".split(" "))
    np.testing.assert_allclose(scores, [0.0, 0.93729472, 0.0], rtol=0, atol=5e-9)
    assert int(np.argmax(scores)) == 1
    raw_is = np.log(3 - 2 + 0.5) - np.log(2 + 0.5)
    assert raw_is < 0 and bm25.idf["is"] == 0.25 * bm25.average_idf
    np.testing.assert_allclose(csr.scores_for_tokens(csr.from_okapi(bm25), ["windy", "London"]),
                               [0.0, 0.93729472, 0.0], rtol=0, atol=5e-9)


def test_filter_mask_semantics():
    srcs = ["CG12", "cg7", "NG1", "PH2", None, "", "XCG1", " ng3"]
    assert retrieval.filter_mask(srcs, "CG, ng").tolist() == [True, True, True, False, False,
                                                             False, False, False]
    assert retrieval.filter_mask(srcs, "cg").tolist() == [True, True, False, False, False, False,
                                                         False, False]


FILTER_SOURCES = ["CG12", "cg7", "NG1", "PH2", None, "", "XCG1", " ng3", "C.1", "CX9", "N(G", 7, "ng"]
# (filter string, frame mask, BM25-style mask): several prefixes are an UN-ESCAPED regex in the
# DataFrame filter (search_engine.py:44-46) and literal prefixes in the BM25 filter (:224-231)
FILTER_CASES = [
    ("CG,NG", [1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1], None),
    ("c., ng", [1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1], [0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1]),
    ("C.", [0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0], None),         # single prefix: literal
    ("CG,", [1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 1], None),         # empty prefix matches every string
    ("N[A-G],PH", [0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1], [0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]),
]


@pytest.mark.parametrize("flt,frame_want,bm25_want", FILTER_CASES)
def test_filter_mask_frame_vs_bm25_semantics(pkg, flt, frame_want, bm25_want):
    bm25_want = frame_want if bm25_want is None else bm25_want
    assert retrieval.filter_mask(FILTER_SOURCES, flt, frame=True).astype(int).tolist() == frame_want
    assert retrieval.filter_mask(FILTER_SOURCES, flt).astype(int).tolist() == bm25_want
    # the product's host-side mask builder follows the same two rules
    assert pkg.engine.prefix_mask(FILTER_SOURCES, flt, frame_semantics=True).astype(int).tolist() == frame_want
    assert pkg.engine.prefix_mask(FILTER_SOURCES, flt).astype(int).tolist() == bm25_want


def test_filter_mask_malformed_regex_raises(pkg):
    import re
    with pytest.raises(re.error):
        retrieval.filter_mask(FILTER_SOURCES, "N(G,CG", frame=True)
    with pytest.raises(re.error):
        pkg.engine.prefix_mask(FILTER_SOURCES, "N(G,CG", frame_semantics=True)
    assert pkg.engine.prefix_mask(FILTER_SOURCES, "N(G,CG").astype(int).tolist() == \
        [1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0]


@pytest.mark.skipif(not reference_loader.available(), reason="reference sources not mounted")
@pytest.mark.parametrize("flt", [c[0] for c in FILTER_CASES] + ["N(G,CG"])
def test_filter_masks_match_reference_filters(pkg, flt):
    """Both filters of the UNMODIFIED reference: the DataFrame one (rows kept by
    _filter_by_filename_type) and the BM25 one (sections that can be returned)."""
    import re
    import pandas as pd
    ref = reference_loader.load_reference()
    se = ref.SearchEngine(None, None)
    df = pd.DataFrame({"source": FILTER_SOURCES})
    try:
        kept = se._filter_by_filename_type(df, flt).index.tolist()
    except re.error:
        with pytest.raises(re.error):
            pkg.engine.prefix_mask(FILTER_SOURCES, flt, frame_semantics=True)
    else:
        assert np.flatnonzero(pkg.engine.prefix_mask(FILTER_SOURCES, flt, frame_semantics=True)).tolist() == kept
        assert np.flatnonzero(retrieval.filter_mask(FILTER_SOURCES, flt, frame=True)).tolist() == kept

    class Sec:
        def __init__(self, source):
            self.metadata = {"source": source}

    class Scores:                                  # bm25.get_scores: every section scores 1.0
        def get_scores(self, _tokens):
            return np.ones(len(strs))

    strs = [s for s in FILTER_SOURCES if isinstance(s, str)]
    ids = [str(i) for i in range(len(strs))]
    got = se._core_bm25_search(["x"], Scores(), [Sec(s) for s in strs], ids, len(strs), flt)
    assert [ids[i] for i in np.flatnonzero(pkg.engine.prefix_mask(strs, flt))] == got


@pytest.mark.skipif(not reference_loader.available(), reason="reference sources not mounted")
def test_restatement_matches_reference_import(small_case):
    """Direct check against the reference's SearchEngine (no golden file in between)."""
    import pandas as pd
    ref = reference_loader.load_reference()
    se = ref.SearchEngine(None, None)
    n = small_case["emb"].shape[0]
    srcs = list(small_case["sources"])
    df = pd.DataFrame({"id": synth.chunk_ids(n, srcs), "document": [""] * n, "source": srcs,
                       "embedding": list(small_case["emb"]), "url": [""] * n})
    for flt in (None, "CG,NG"):
        mask = None if flt is None else retrieval.filter_mask(srcs, flt)
        for q in range(4):
            res = se.similarity_search_with_embedding(small_case["queries"][q], df,
                                                      "voyage-3-large", 10, flt)
            rows, scores = retrieval.dense_topk(small_case["queries"][q], small_case["emb"], 10, mask)
            full = retrieval.dense_scores(small_case["queries"][q], small_case["emb"])
            check_topk(rows, scores, res.index.to_numpy(), res["similarity"].to_numpy(), full,
                       f"ref dense q{q}")
    lists = [(["a", "b", "c", "d"], "voyage-3-large"), (["c", "x", "a"], "BM25"),
             (["x", "y"], "other")]
    assert se.weighted_reciprocal_rank_fusion(lists, WEIGHTS, 40) == \
        retrieval.weighted_rrf(lists, WEIGHTS, 40)


@pytest.mark.skipif(not reference_loader.available(), reason="reference sources not mounted")
def test_orchestrator_restatement_matches_reference_method(small_case, built):
    """oracle.orchestrator.retrieve_documents == the UNMODIFIED
    RetrievalEvaluationSystem.retrieve_documents (query_rag_retrieval.py:149-411) driven with the
    reference's own SearchEngine: several dense models, BM25, filters, full ranking, return_docs."""
    import types
    import pandas as pd
    from oracle import orchestrator
    ref = reference_loader.load_reference()
    n = small_case["emb"].shape[0]
    srcs = list(small_case["sources"])
    ids = synth.chunk_ids(n, srcs)
    emb = small_case["emb"]
    rng = np.random.default_rng(3)
    emb2 = (emb + 0.3 * rng.standard_normal(emb.shape)).astype(np.float32)   # a second "model"

    def frame(e):
        return pd.DataFrame({"id": ids, "document": [f"doc {i}" for i in range(n)], "source": srcs,
                             "embedding": list(e), "url": [""] * n})
    ix, okapi = built
    sections = [types.SimpleNamespace(page_content=f"doc {i}", metadata={"id": ids[i], "source": srcs[i]})
                for i in range(n)]
    system = object.__new__(ref.RetrievalEvaluationSystem)     # no Voyage client, no databases on disk
    system.config = ref.Config()
    system.search_engine = ref.SearchEngine(None, None)
    nice = ref.InfoSource("nice")
    system.embeddings_data = {nice: {"voyage-3-large": frame(emb), "Qwen3": frame(emb2)}}
    system.bm25_data = {nice: (okapi, sections, ids)}
    weights = {"voyage-3-large": 5.0, "Qwen3": 2.0, "BM25": 1.0}
    cases = [dict(similarity_k=10, common_sections_n=10, use_hybrid_search=True),
             dict(similarity_k=10, common_sections_n=7, use_hybrid_search=False),
             dict(similarity_k=25, common_sections_n=15, use_hybrid_search=True,
                  filename_type_filter="CG, NG"),
             dict(similarity_k=3000, common_sections_n=40, use_hybrid_search=True),   # k >= N
             dict(similarity_k=10, common_sections_n=10, use_hybrid_search=True, return_docs=True),
             dict(similarity_k=10, common_sections_n=10, use_hybrid_search=True,
                  filename_type_filter="ZZ")]
    for kw in cases:
        for q in range(3):
            qe = {"voyage-3-large": small_case["queries"][q], "Qwen3": small_case["queries"][q + 3]}
            toks = synth.token_strings(small_case["term_queries"][q])
            want = system.retrieve_documents(qe, None, toks, info_source="NICE", model_weights=weights,
                                             wrrf_k=40, use_reranker=False, **kw)
            got = orchestrator.retrieve_documents(system, nice, qe, None, toks, model_weights=weights,
                                                  wrrf_k=40, use_reranker=False, **kw)
            if kw.get("return_docs"):
                assert [d["id"] for d in got] == [d["id"] for d in want]
                assert [d["similarity"] for d in got] == [d["similarity"] for d in want]
            else:
                assert got == want, (kw, q)
            assert len(want) > 0 or kw.get("filename_type_filter") == "ZZ"


@pytest.mark.skipif(not reference_loader.available(), reason="reference sources not mounted")
def test_wrrf_and_topk_restatements_fuzzed_against_the_reference():
    """Random ranked lists (overlaps, empty lists, duplicate ids inside a list, unknown model
    names, zero / negative weights, the three wrrf_k values the callers use) and random score
    vectors with planted ties: bit-identical fusion output and identical filtered BM25 order."""
    ref = reference_loader.load_reference()
    se = ref.SearchEngine(None, None)
    rng = np.random.default_rng(20261018)
    names = ["voyage-3-large", "BM25", "Qwen3", "unlisted"]
    weights = {"voyage-3-large": 5.0, "BM25": 1.0, "Qwen3": 0.0}
    for trial in range(200):
        lists = []
        for _ in range(int(rng.integers(1, 6))):
            n = int(rng.integers(0, 40))
            ids = [f"id{int(i)}" for i in rng.integers(0, 60, size=n)]
            lists.append((ids, names[int(rng.integers(0, len(names)))]))
        w = dict(weights)
        if trial % 7 == 0:
            w["BM25"] = -0.5
        k = (40, 50, 60)[trial % 3]
        assert se.weighted_reciprocal_rank_fusion(lists, w, k) == retrieval.weighted_rrf(lists, w, k)

    class Sec:
        def __init__(self, source):
            self.metadata = {"source": source}

    for trial in range(50):
        n = int(rng.integers(1, 200))
        scores = np.round(rng.random(n) * 4, 1)                  # many exact ties
        srcs = [("CG", "NG", "PH", "cg")[int(i)] + str(trial) for i in rng.integers(0, 4, size=n)]
        ids = [str(i) for i in range(n)]

        class Okapi:
            def get_scores(self, _tokens, s=scores):
                return s

        for flt in ("CG", "cg, ng", "ZZ"):
            k = int(rng.integers(1, n + 5))
            got = se._core_bm25_search(["t"], Okapi(), [Sec(s) for s in srcs], ids, k, flt)
            want = retrieval.bm25_topk(scores, k, srcs, flt)
            assert got == [ids[i] for i in want]


@pytest.mark.skipif(not reference_loader.available(), reason="reference sources not mounted")
def test_committed_golden_is_what_the_reference_produces_today(small_case, tmp_path):
    """Re-runs oracle/make_golden.py's small case through the UNMODIFIED reference (SQLite chunks
    table + BM25 pickle on disk -> DatabaseManager -> SearchEngine) and requires every array of
    tests/golden/small_case.npz to come out again bit for bit."""
    from oracle import make_golden
    inputs = make_golden.small_case_inputs()
    out = make_golden.run_reference(inputs, ks=(10, 100, 3000), filters=(None, "CG,NG", "cg", "ZZ"),
                                    tmpdir=str(tmp_path))
    fresh = {**inputs, **out}
    assert set(fresh) == set(small_case)
    for name, want in small_case.items():
        got = np.asarray(fresh[name])
        assert got.shape == want.shape, name
        if want.dtype == object:
            assert got.tolist() == want.tolist(), name
        else:
            np.testing.assert_array_equal(got, want, err_msg=name)


@pytest.mark.skipif(not reference_loader.available(), reason="reference sources not mounted")
def test_committed_config0_golden_is_what_the_reference_produces_today(config0_golden, tmp_path):
    """BASELINE configs[0] (20k chunks x 1024-d, 100 queries, filters None and "CG, NG") through the
    UNMODIFIED reference again: every stored array of tests/golden/config0_outputs.npz comes out
    bit for bit (the GPU suite checks the drop-in classes against exactly these arrays)."""
    from oracle import make_golden
    case = make_golden.config0_inputs()
    np.testing.assert_allclose(make_golden.checksum(case), config0_golden["input_checksum"], rtol=1e-12)
    out = make_golden.run_reference(case, ks=(10,), filters=(None, "CG, NG"), tmpdir=str(tmp_path))
    out.pop("bm25_all_scores")
    assert set(out) | {"input_checksum"} == set(config0_golden)
    for name, got in out.items():
        want = config0_golden[name]
        got = np.asarray(got)
        assert got.shape == want.shape, name
        if want.dtype == object:
            assert got.tolist() == want.tolist(), name
        else:
            np.testing.assert_array_equal(got, want, err_msg=name)


# ---------------------------------------------------------------------------------------
# the slab-wise / shard-wise forms used at 10M+ rows equal the whole-corpus statements
# ---------------------------------------------------------------------------------------
def test_slab_and_shard_forms_equal_the_whole_corpus_oracle(small_case):
    from oracle import pipeline, slabs
    case = small_case
    ix, _ = csr_from_case(case)
    emb, queries = case["emb"], case["queries"]
    n = emb.shape[0]
    bounds = [0, 300, 1100, n]
    pieces = [(bounds[i], emb[bounds[i]:bounds[i + 1]]) for i in range(3)]
    ids, sc = slabs.dense_topk_slabs(iter(pieces), queries[:4], 10)
    for q in range(4):
        w_ids, w_sc = retrieval.dense_topk(queries[q], emb, 10)
        retrieval.assert_ranking_matches(ids[q], sc[q], w_ids, w_sc,
                                         all_scores=retrieval.dense_scores(queries[q], emb))
    fetch = lambda t: (ix.post_doc[ix.term_ptr[t]:ix.term_ptr[t + 1]],       # noqa: E731
                       ix.post_tf[ix.term_ptr[t]:ix.term_ptr[t + 1]])
    weights = {"voyage-3-large": 5.0, "BM25": 1.0}
    for q in range(4):
        terms = [int(t) if t < ix.idf.shape[0] else -1 for t in case["term_queries"][q]]
        terms = terms + terms[:1] + [-1]          # a duplicate and an unknown term
        want = csr.scores(ix, terms)
        got = slabs.bm25_scores_subindex(fetch, terms, ix.doc_len, lambda t: float(ix.idf[t]),
                                         ix.avgdl, ix.k1, ix.b)
        assert np.array_equal(got, want)
        # shards: the same documents' postings with GLOBAL idf / avgdl
        shards = []
        for i in range(3):
            lo, hi = bounds[i], bounds[i + 1]
            local = csr.from_token_ids(case["doc_ptr"][lo:hi + 1] - case["doc_ptr"][lo],
                                       case["tokens"][case["doc_ptr"][lo]:case["doc_ptr"][hi]],
                                       int(case["vocab"]), ix.k1, ix.b, 0.05)
            local.idf, local.avgdl = ix.idf, ix.avgdl
            shards.append((lo, emb[lo:hi], local))
        whole = pipeline.hybrid_query(queries[q], emb, ix, terms, 10, 10, weights, 40.0, 10)
        parts = pipeline.hybrid_query_sharded(queries[q], shards, terms, 10, 10, weights, 40.0, 10)
        assert np.allclose(parts["bm25_all"], whole["bm25_all"], rtol=1e-12, atol=0)
        retrieval.assert_ranking_matches(parts["dense_ids"], parts["dense_scores"],
                                         whole["dense_ids"], whole["dense_scores"],
                                         all_scores=retrieval.dense_scores(queries[q], emb))
        assert [i for i, _ in parts["fused"]] == [i for i, _ in whole["fused"]]


@pytest.mark.skipif(not reference_loader.available(), reason="reference sources not mounted")
def test_cpu_search_engine_equals_the_reference_class(small_case):
    """oracle.cpu_search_engine (the CPU arm of the evaluator-shaped timing on the GPU box) against
    the reference's SearchEngine imported verbatim, on the same frames / BM25 object."""
    import types
    import pandas as pd
    from oracle import cpu_search_engine
    ref_cls = reference_loader.load_reference().SearchEngine
    ref, mine = ref_cls(None, None), cpu_search_engine.CpuSearchEngine()
    case = small_case
    n = case["emb"].shape[0]
    srcs = list(case["sources"])
    ids = synth.chunk_ids(n, srcs)
    df = pd.DataFrame({"id": ids, "document": [""] * n, "source": srcs, "embedding": list(case["emb"]),
                       "url": [""] * n})
    okapi = okapi_from_case(case)
    sections = [types.SimpleNamespace(page_content="", metadata={"id": ids[i], "source": srcs[i]})
                for i in range(n)]
    for flt in (None, "CG, NG", "ZZ"):
        for q in range(4):
            a = ref.similarity_search_with_embedding(case["queries"][q], df, "m", 25, flt)
            b = mine.similarity_search_with_embedding(case["queries"][q], df, "m", 25, flt)
            assert a["id"].tolist() == b["id"].tolist()
            if len(a):
                assert np.array_equal(a["similarity"].to_numpy(), b["similarity"].to_numpy())
            toks = synth.token_strings(case["term_queries"][q])
            assert (ref.bm25_search_preprocessed(toks, okapi, sections, ids, 25, flt)
                    == mine.bm25_search_preprocessed(toks, okapi, sections, ids, 25, flt))
    lists = [(ids[:30], "voyage-3-large"), (ids[10:50][::-1], "BM25")]
    w = {"voyage-3-large": 5.0, "BM25": 1.0}
    assert ref.weighted_reciprocal_rank_fusion(lists, w, 40) == mine.weighted_reciprocal_rank_fusion(lists, w, 40)


# ---------------------------------------------------------------------------------------
# The candidate-driven BM25 top-k (csrc/anr_bm25_ms.cu): its pruning decisions, restated in
# oracle/maxscore_model.py, never lose a top-k document
def test_maxscore_model_keeps_the_exact_topk():
    from oracle import maxscore_model as mm
    synth_mod = synth
    n, vocab = 12_000, 2500
    doc_ptr, tokens = synth_mod.zipf_corpus(n, vocab, 1.1, seed=5, len_lo=40, len_hi=120)
    ix = csr.from_token_ids(doc_ptr, tokens, vocab, 1.7, 0.83, 0.05)
    w = mm.posting_weights(ix)
    idf32 = ix.idf.astype(np.float32)

    def exhaustive(terms, k, allowed):
        acc = np.zeros(n, np.float32)
        for t in terms:                       # fp32 fma in query order, as the kernel sums
            if 0 <= t < vocab and idf32[t] > 0:
                a, b = int(ix.term_ptr[t]), int(ix.term_ptr[t + 1])
                d = ix.post_doc[a:b]
                acc[d] = (np.float64(idf32[t]) * w[a:b].astype(np.float64) + acc[d]).astype(np.float32)
        if allowed is not None:
            acc = np.where(allowed, acc, np.float32(-1))
        order = np.lexsort((np.arange(n), -acc))[:k]
        return [(float(acc[i]), int(i)) for i in order if acc[i] > 0]

    tq = synth_mod.zipf_queries(6, 8, vocab, 1.1, seed=6)
    allowed = np.random.default_rng(1).random(n) < 0.5
    df = np.diff(ix.term_ptr)
    rare = int(np.argmin(np.where(df > 0, df, 1 << 30)))      # the shortest posting list
    k_rare = int(df[rare]) + 1                                # more than the documents that match
    cases = [(list(map(int, t)), 10, None) for t in tq]
    cases += [(list(map(int, tq[0])), 100, None), (list(map(int, tq[1])), 10, allowed),
              ([int(tq[2][0])] * 3 + [-1, vocab + 5], 10, None), ([0, 1, 2], 10, None),
              ([rare], k_rare, None), ([], 10, None)]
    flagged_cases = 0
    for terms, k, al in cases:
        flagged, surv, stats = mm.topk(ix, w, terms, k, al)
        want = exhaustive(terms, k, al)
        if flagged:                           # handed to the exhaustive scan: nothing to prove
            flagged_cases += 1
            assert len(want) < k or stats["survivors"] > mm.SURVIVORS, (terms, k, stats)
            continue
        assert surv[:k] == want, (terms, k, stats)
        # the pruning is real: far fewer postings are touched than the query's lists hold
        if terms and k == 10 and al is None and len(terms) == 8:
            sum_df = sum(int(df[t]) for t in terms if 0 <= t < vocab)
            assert stats["streamed"] + stats["sampled"] < sum_df
    assert flagged_cases == 1                 # the rare term alone: fewer than k matching documents
