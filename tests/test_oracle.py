"""CPU tests: the oracles (numpy + C) against the golden fixtures generated from the reference itself
(oracle/make_golden.py) and against the reference's own known-answer vectors."""
import numpy as np
import pytest

from _check import check_topk_rows, lists_from_flat, load_golden
from oracle import c_oracle as CO
from oracle import oracle as O

METHODS = O.METHODS


def key(m):
    return m.replace("+", "plus")


def test_known_answer_vectors():
    """tests/core/test_scoring.py:26-98, tests/core/test_topk.py:16-59, tests/core/test_selection.py:34-58."""
    g = load_golden("known_answers")
    data, indices, indptr = np.array([1.0, 2.0], np.float32), np.array([0, 1], np.int32), np.array([0, 1, 2], np.int64)
    np.testing.assert_array_equal(O.score_dense(data, indptr, indices, 3, [0, 1]), np.array([1.0, 2.0, 0.0], np.float32))
    np.testing.assert_array_equal(O.score_dense(data, indptr, indices, 3, [0, 1]), g["ka1_jit"])
    np.testing.assert_array_equal(O.score_dense(data, indptr, indices, 3, [0, 1]), g["ka1_legacy"])
    np.testing.assert_array_equal(O.score_dense(data, indptr, indices, 3, []), g["ka1_empty"])
    np.testing.assert_array_equal(CO.scores_dense(data, indices, indptr, 3, [0, 1]), g["ka1_jit"])
    d2, i2 = np.array([1.5, 2.5], np.float32), np.array([0, 0], np.int32)
    np.testing.assert_array_equal(O.score_dense(d2, indptr, i2, 2, [0, 1]), np.array([4.0, 0.0], np.float32))
    np.testing.assert_array_equal(CO.scores_dense(d2, i2, indptr, 2, [0, 1]), g["ka2_jit"])
    s, i = O.topk(g["topk_scores_in"], 5)
    np.testing.assert_array_equal(s, g["topk_s"])
    np.testing.assert_array_equal(i, g["topk_i"])
    np.testing.assert_array_equal(i, np.argsort(g["topk_scores_in"])[-5:][::-1])
    s, i = CO.topk(g["topk_scores_in"].astype(np.float32), 5)
    np.testing.assert_array_equal(i, g["topk_i"])
    s, i = O.topk(np.array([3.0, 1.0, 4.0, 2.0], np.float32), 2)
    np.testing.assert_array_equal(s, g["sel_s"])
    np.testing.assert_array_equal(i, g["sel_i"])
    np.testing.assert_array_equal(i, [2, 0])


def test_readme_corpus_golden_vectors():
    """SURVEY.md A.6 values and the reference run of the README corpus."""
    g = load_golden("readme_corpus")
    expect = {
        "robertson": [0.7448772192, 0.3109349906, 0.3389191329, 0.3389191329],
        "lucene": [1.0584375858, 0.4418248832, 0.4815891385, 0.4815891385],
        "atire": [3.0468008518, 1.2718297243, 1.3862943649, 1.3862943649],
        "bm25l": [5.4345993996, 4.4464902878, 4.5148978233, 4.5148978233],
        "bm25+": [7.5608210564, 5.5001435280, 5.6330327988, 5.6330327988],
    }
    corpus = lists_from_flat(g["corpus_tokens"], g["corpus_ptr"])
    q = g["query_ids"].tolist()
    V = int(g["n_vocab"])
    for m in METHODS:
        np.testing.assert_allclose(g[f"{key(m)}_dense"], np.array(expect[m], np.float32), rtol=1e-7)
        data, indices, indptr, nonocc = O.build_index(corpus, V, m)
        np.testing.assert_array_equal(data, g[f"{key(m)}_data"])
        np.testing.assert_array_equal(indices, g[f"{key(m)}_indices"])
        np.testing.assert_array_equal(indptr, g[f"{key(m)}_indptr"])
        if nonocc is not None:
            np.testing.assert_array_equal(nonocc, g[f"{key(m)}_nonocc"])
        np.testing.assert_array_equal(O.get_scores(data, indptr, indices, 4, q, nonocc), g[f"{key(m)}_dense"])
        for mk, mask in (("f32", np.array([1, 0, 0, 1], np.float32)), ("i32", np.array([1, 0, 0, 1], np.int32)),
                         ("bool", np.array([1, 0, 0, 1], np.bool_))):
            np.testing.assert_array_equal(O.get_scores(data, indptr, indices, 4, q, nonocc, mask), g[f"{key(m)}_dense_mask_{mk}"])
            sh = None if nonocc is None else O.query_shift(nonocc, q)
            np.testing.assert_array_equal(CO.scores_dense(data, indices, indptr, 4, q, mask=mask.astype(np.float64), shift=sh),
                                          g[f"{key(m)}_dense_mask_{mk}"])
    np.testing.assert_array_equal(g["lucene_indptr"], np.arange(21))
    np.testing.assert_array_equal(g["lucene_indices"], [0] * 4 + [1] * 6 + [2] * 5 + [3] * 5)


@pytest.mark.parametrize("name", ["synth_small", "synth_medium"])
def test_synthetic_fixtures(name):
    g = load_golden(name)
    corpus = lists_from_flat(g["corpus_tokens"], g["corpus_ptr"])
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    V, k, N = int(g["n_vocab"]), int(g["k"]), len(g["corpus_ptr"]) - 1
    for m in METHODS:
        data, indices, indptr, nonocc = O.build_index(corpus, V, m)
        np.testing.assert_array_equal(data, g[f"{key(m)}_data"])
        np.testing.assert_array_equal(indices, g[f"{key(m)}_indices"])
        np.testing.assert_array_equal(indptr, g[f"{key(m)}_indptr"])
        if nonocc is not None:
            np.testing.assert_array_equal(nonocc, g[f"{key(m)}_nonocc"])
        ids, sc = O.retrieve(data, indptr, indices, N, queries, k, nonocc)
        np.testing.assert_array_equal(sc, g[f"{key(m)}_scores"])
        np.testing.assert_array_equal(ids, g[f"{key(m)}_ids"])
        nd = g[f"{key(m)}_dense"].shape[0]
        for qi in range(nd):
            np.testing.assert_array_equal(O.get_scores(data, indptr, indices, N, queries[qi], nonocc), g[f"{key(m)}_dense"][qi])
            np.testing.assert_array_equal(O.get_scores(data, indptr, indices, N, queries[qi], nonocc, g["maskf64"]),
                                          g[f"{key(m)}_dense_maskf64"][qi])
        # C oracle: scores bit-exact, ids tie-aware
        sh = None if nonocc is None else np.array([O.query_shift(nonocc, q) for q in queries], np.float32)
        dense_fn = lambda qi: O.get_scores(data, indptr, indices, N, queries[qi], nonocc)  # noqa: E731
        ids2, sc2 = CO.retrieve(data, indices, indptr, N, g["query_tokens"], g["query_ptr"], k, shift=sh)
        check_topk_rows(ids2, sc2, g[f"{key(m)}_ids"], g[f"{key(m)}_scores"], dense_fn, what=f"C {name} {m}")
        for mk in ("mask01", "maskf64"):
            ids3, sc3 = CO.retrieve(data, indices, indptr, N, g["query_tokens"], g["query_ptr"], k, shift=sh, mask=g[mk])
            np.testing.assert_array_equal(sc3, g[f"{key(m)}_scores_{mk}"])
    d, _, _, _ = O.build_index(corpus, V, "atire", k1=1.2, b=0.6)
    np.testing.assert_array_equal(d, g["atire_k12_data"])
    d, _, _, n = O.build_index(corpus, V, "bm25+", k1=0.9, b=0.4, delta=1.0)
    np.testing.assert_array_equal(d, g["bm25plus_k09_data"])
    np.testing.assert_array_equal(n, g["bm25plus_k09_nonocc"])
    d, _, _, _ = O.build_index(corpus, V, "lucene", idf_method="robertson")
    np.testing.assert_array_equal(d, g["lucene_idfrobertson_data"])


def test_c_oracle_threads_agree():
    g = load_golden("synth_medium")
    N, k = len(g["corpus_ptr"]) - 1, int(g["k"])
    a = CO.retrieve(g["lucene_data"], g["lucene_indices"], g["lucene_indptr"], N, g["query_tokens"], g["query_ptr"], k, n_threads=1)
    b = CO.retrieve(g["lucene_data"], g["lucene_indices"], g["lucene_indptr"], N, g["query_tokens"], g["query_ptr"], k, n_threads=4)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
