"""GPU index build (bm25cu_index_build) against the reference fixtures and the numpy oracle:
indices / indptr bit-exact, data bit-exact (the north star allows 1e-5 relative; the build reproduces every float)."""
import numpy as np
import pytest

from _check import check_topk_rows, lists_from_flat, load_golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu
METHODS = O.METHODS


def key(m):
    return m.replace("+", "plus")


def _build(corpus_tokens, corpus_ptr, n_vocab, method, idf_method=None, k1=1.5, b=0.75, delta=0.5):
    from bm25s_b200 import DeviceIndex

    dev = DeviceIndex.build(corpus_tokens, corpus_ptr, n_vocab, method, idf_method or method, k1, b, delta)
    out = dev.download()
    dev.close()
    return out


@pytest.mark.parametrize("name", ["readme_corpus", "synth_small", "synth_medium"])
def test_build_matches_reference_fixture(name):
    g = load_golden(name)
    V = int(g["n_vocab"])
    for m in METHODS:
        data, indices, indptr, nonocc = _build(g["corpus_tokens"], g["corpus_ptr"], V, m)
        np.testing.assert_array_equal(indptr, g[f"{key(m)}_indptr"])
        np.testing.assert_array_equal(indices, g[f"{key(m)}_indices"])
        np.testing.assert_array_equal(data, g[f"{key(m)}_data"])
        if m in ("bm25l", "bm25+"):
            np.testing.assert_array_equal(nonocc, g[f"{key(m)}_nonocc"])
        else:
            assert nonocc is None
    if name != "readme_corpus":
        d = _build(g["corpus_tokens"], g["corpus_ptr"], V, "atire", k1=1.2, b=0.6)[0]
        np.testing.assert_array_equal(d, g["atire_k12_data"])
        d, _, _, n = _build(g["corpus_tokens"], g["corpus_ptr"], V, "bm25+", k1=0.9, b=0.4, delta=1.0)
        np.testing.assert_array_equal(d, g["bm25plus_k09_data"])
        np.testing.assert_array_equal(n, g["bm25plus_k09_nonocc"])
        d = _build(g["corpus_tokens"], g["corpus_ptr"], V, "lucene", idf_method="robertson")[0]
        np.testing.assert_array_equal(d, g["lucene_idfrobertson_data"])


def test_build_edge_cases_vs_oracle():
    """Empty documents, long documents (> one sort tile), heavy repetition, unused vocabulary ids, vocab > 2^16."""
    rng = np.random.default_rng(11)
    n_vocab = 70_000
    docs = []
    for i in range(400):
        r = i % 8
        if r == 0:
            docs.append([])
        elif r == 1:
            docs.append([7] * int(rng.integers(1, 50)))
        elif r == 2:
            docs.append(rng.integers(0, n_vocab, int(rng.integers(3000, 9000))).tolist())
        else:
            docs.append(rng.integers(0, 300, int(rng.integers(1, 200))).tolist())
    ptr = np.zeros(len(docs) + 1, np.int64)
    np.cumsum([len(d) for d in docs], out=ptr[1:])
    tok = np.array([t for d in docs for t in d], np.int32)
    for m in ("lucene", "bm25l", "bm25+", "atire", "robertson"):
        data, indices, indptr, nonocc = _build(tok, ptr, n_vocab, m)
        rd, ri, rp, rn = O.build_index(docs, n_vocab, m)
        np.testing.assert_array_equal(indptr, rp)
        np.testing.assert_array_equal(indices, ri)
        np.testing.assert_array_equal(data, rd)
        if rn is not None:
            np.testing.assert_array_equal(nonocc, rn)
    with pytest.raises(ValueError):
        _build(np.array([0, n_vocab], np.int32), np.array([0, 2], np.int64), n_vocab, "lucene")


def test_bm25_class_index_retrieve_save_load(tmp_path):
    """BM25(backend="cuda").index -> retrieve == the reference fixture; save/load round trip keeps the arrays."""
    from bm25s_b200 import BM25

    g = load_golden("synth_small")
    corpus = lists_from_flat(g["corpus_tokens"], g["corpus_ptr"])
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    V, k, N = int(g["n_vocab"]), int(g["k"]), len(corpus)
    for m in METHODS:
        r = BM25(method=m, backend="cuda")
        r.index((corpus, {i: i for i in range(V)}), create_empty_token=False, show_progress=False)
        np.testing.assert_array_equal(r.scores["data"], g[f"{key(m)}_data"])
        np.testing.assert_array_equal(r.scores["indices"], g[f"{key(m)}_indices"])
        np.testing.assert_array_equal(r.scores["indptr"], g[f"{key(m)}_indptr"])
        res = r.retrieve(queries, k=k, show_progress=False)
        assert res.documents.dtype == np.int64 and res.scores.dtype == np.float32
        data, indices, indptr = r.scores["data"], r.scores["indices"], r.scores["indptr"]
        dense_fn = lambda qi: O.get_scores(data, indptr, indices, N, queries[qi], r.nonoccurrence_array)  # noqa: E731
        check_topk_rows(res.documents, res.scores, g[f"{key(m)}_ids"], g[f"{key(m)}_scores"], dense_fn, what=f"BM25 {m}")
        np.testing.assert_array_equal(r.get_scores(queries[0]), g[f"{key(m)}_dense"][0])
        res_m = r.retrieve(queries, k=k, weight_mask=g["mask01"])
        np.testing.assert_array_equal(res_m.scores, g[f"{key(m)}_scores_mask01"])
        d = tmp_path / f"idx_{key(m)}"
        r.save(d)
        r2 = BM25.load(d, mmap=(m == "lucene"))
        np.testing.assert_array_equal(r2.scores["data"], data)
        res2 = r2.retrieve(queries, k=k)
        np.testing.assert_array_equal(res2.scores, res.scores)
        np.testing.assert_array_equal(res2.documents, res.documents)
