"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/bm25cu.h declares, the
argument checks that need no GPU, the reference-API mirror (errors, containers, file formats)."""
import json
import os
import re

import numpy as np
import pytest

from _check import lists_from_flat, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from bm25s_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "bm25cu.h")).read()
    declared = set(re.findall(r"\b(bm25cu_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/bm25cu.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert L.bm25cu_version() >= 100
    assert _lib.device_count() >= 0


def test_abi_argument_checks_without_gpu():
    import ctypes as C

    from bm25s_b200 import _lib

    L = _lib.lib()
    assert L.bm25cu_index_sizes(None, None, None, None, None, None) == 1
    assert "null handle" in _lib.last_error()
    assert L.bm25cu_retrieve(None, None, None, 0, 0, 1, None, None, None, None, None) == 1
    h = C.c_void_p()
    ip = np.array([0, 1], np.int64)
    # indptr that does not span nnz is rejected before any CUDA call
    rc = L.bm25cu_index_upload(None, None, _lib.ptr(ip, C.c_int64), C.c_int64(5), 1, 3, None, 0, 0, 3, C.byref(h))
    assert rc == 1
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert L.bm25cu_topk_dense(None, 0, C.c_int64(3), 5, 1, None, None, 0) == 1  # k > n


def _retriever_with_scores():
    from bm25s_b200 import BM25

    g = load_golden("synth_small")
    r = BM25(method="bm25+")
    r.scores = {"data": g["bm25plus_data"], "indices": g["bm25plus_indices"], "indptr": g["bm25plus_indptr"],
                "num_docs": len(g["corpus_ptr"]) - 1}
    r.nonoccurrence_array = g["bm25plus_nonocc"]
    r.vocab_dict = {i: i for i in range(int(g["n_vocab"]))}
    r.unique_token_ids_set = set(r.vocab_dict.values())
    return r, g


def test_retrieve_validation_errors_match_reference():
    """bm25s/__init__.py:759-845 (tests/core/test_retrieve.py:135-186, tests/core/test_core_coverage.py:276-313)."""
    r, g = _retriever_with_scores()
    n = r.scores["num_docs"]
    with pytest.raises(ValueError, match="Please set with a smaller k or increase the size of corpus."):
        r.retrieve([[1, 2]], k=n + 1)
    with pytest.raises(ValueError, match="return_as"):
        r.retrieve([[1, 2]], k=1, return_as="nope")
    with pytest.raises(ValueError, match="weight_mask must be a numpy array"):
        r.retrieve([[1, 2]], k=1, weight_mask=[1] * n)
    with pytest.raises(ValueError, match="1D"):
        r.retrieve([[1, 2]], k=1, weight_mask=np.ones((n, 1)))
    with pytest.raises(ValueError, match="same as the length of the corpus"):
        r.retrieve([[1, 2]], k=1, weight_mask=np.ones(n + 1))
    with pytest.raises(ValueError):
        r.retrieve(({"a": 0}, [[0]]), k=1)  # (vocab, ids) swapped
    with pytest.raises(ValueError):
        r.retrieve(([[0]], [[0]]), k=1)  # ids twice
    with pytest.raises(ValueError):
        r.retrieve(({"a": 0},), k=1)
    # all-OOV integer query without an empty token
    with pytest.raises(ValueError, match="does not contain any tokens that are in the vocabulary"):
        r.retrieve([[10 ** 6]], k=1)
    # id >= vocabulary size
    with pytest.raises(ValueError, match="maximum token ID"):
        r.retrieve_flat(np.array([10 ** 6]), np.array([0, 1]), k=1)
    del r.unique_token_ids_set
    with pytest.raises(ValueError, match="unique_token_ids_set"):
        r.retrieve([[1, 2]], k=1)
    # k = 0 returns (Q, 0) without touching the device
    r.unique_token_ids_set = set(r.vocab_dict.values())
    res = r.retrieve([[1, 2], [3]], k=0)
    assert res.documents.shape == (2, 0) and res.scores.shape == (2, 0)


def test_backend_must_be_cuda():
    from bm25s_b200 import BM25

    with pytest.raises(ValueError, match="only implements backend='cuda'"):
        BM25(backend="numpy")
    assert BM25(backend="auto").backend == "cuda"


def test_query_shift_is_numpy_bit_exact():
    r, g = _retriever_with_scores()
    from bm25s_b200.bm25 import _flatten

    rng = np.random.default_rng(0)
    queries = [rng.integers(0, int(g["n_vocab"]), int(n)).tolist() for n in rng.integers(0, 40, 200)]
    flat, ptr = _flatten(queries)
    shift = r._query_shifts(flat.astype(np.int32), ptr)
    for qi, q in enumerate(queries):
        ref = r.nonoccurrence_array[np.asarray(q, np.int32)].sum() if len(q) else np.float32(0)
        assert shift[qi] == ref, (qi, len(q))


def test_flatten_layout():
    from bm25s_b200.bm25 import _flatten

    flat, ptr = _flatten([[3, 1], [], [7]])
    np.testing.assert_array_equal(flat, [3, 1, 7])
    np.testing.assert_array_equal(ptr, [0, 2, 2, 3])
    flat, ptr = _flatten([])
    assert len(flat) == 0 and ptr.tolist() == [0]


def test_save_load_file_format(tmp_path):
    """Same file names / params keys as the reference (bm25s/__init__.py:945-951, :1027-1040); loads without a GPU."""
    from bm25s_b200 import BM25

    r, g = _retriever_with_scores()
    r.vocab_dict = {"café": 0, "b": 1, "": 2}
    corpus = ["café au lait", {"id": "x", "text": "b"}] + ["d"] * (r.scores["num_docs"] - 2)
    r.save(tmp_path, corpus=corpus)
    for f in ("data.csc.index.npy", "indices.csc.index.npy", "indptr.csc.index.npy", "vocab.index.json",
              "params.index.json", "nonoccurrence_array.index.npy", "corpus.jsonl", "corpus.mmindex.json"):
        assert (tmp_path / f).exists(), f
    params = json.load(open(tmp_path / "params.index.json"))
    assert set(params) == {"k1", "b", "delta", "method", "idf_method", "dtype", "int_dtype", "num_docs", "version", "backend"}
    for mmap in (False, True):
        r2 = BM25.load(tmp_path, load_corpus=True, mmap=mmap)
        assert r2.backend == "cuda" and r2.method == "bm25+"
        np.testing.assert_array_equal(r2.scores["data"], r.scores["data"])
        np.testing.assert_array_equal(r2.nonoccurrence_array, r.nonoccurrence_array)
        assert r2.vocab_dict["café"] == 0
        assert r2.corpus[0] == {"id": 0, "text": "café au lait"} and r2.corpus[1]["id"] == "x"
        assert len(r2.corpus) == r.scores["num_docs"]
    # an index written by the reference records backend "numpy"/"numba": it loads as cuda
    params["backend"] = "numba"
    json.dump(params, open(tmp_path / "params.index.json", "w"))
    assert BM25.load(tmp_path).backend == "cuda"
    os.remove(tmp_path / "nonoccurrence_array.index.npy")
    with pytest.raises(FileNotFoundError):
        BM25.load(tmp_path)


def test_tokenized_round_trip():
    from bm25s_b200.tokenization import Tokenized, convert_tokenized_to_string_list

    t = Tokenized(ids=[[0, 1], [1]], vocab={"a": 0, "b": 1})
    assert convert_tokenized_to_string_list(t) == [["a", "b"], ["b"]]


def test_tokenized_remap_equals_string_round_trip():
    """the one-table remap of Tokenized queries gives what the reference's id -> str -> id path gives
    (tokenization.py:513-521 then bm25s/__init__.py:572-579): OOV dropped, order and duplicates kept."""
    from bm25s_b200.bm25 import _flatten
    from bm25s_b200.tokenization import Tokenized, convert_tokenized_to_string_list

    r, g = _retriever_with_scores()
    rng = np.random.default_rng(3)
    words = [f"w{i}" for i in range(60)]
    r.vocab_dict = {w: i for i, w in enumerate(words[:40])}       # index vocabulary: w0..w39
    qvocab = {w: j for j, w in enumerate(rng.permutation(words[20:]).tolist())}  # query vocabulary: w20..w59, shuffled ids
    ids = [rng.integers(0, len(qvocab), int(n)).tolist() for n in rng.integers(0, 12, 300)]
    tok = Tokenized(ids=ids, vocab=qvocab)
    flat, ptr = r._remap_tokenized(tok)
    ref = [r.get_tokens_ids(q) for q in convert_tokenized_to_string_list(tok)]
    ref_flat, ref_ptr = _flatten(ref)
    np.testing.assert_array_equal(flat, ref_flat)
    np.testing.assert_array_equal(ptr, ref_ptr)
    assert any(len(q) == 0 for q in ref) and sum(len(q) for q in ref) < sum(len(q) for q in ids)  # OOV was exercised
    with pytest.raises(KeyError):
        r._remap_tokenized(Tokenized(ids=[[len(qvocab) + 5]], vocab=qvocab))
    flat, ptr = r._remap_tokenized(Tokenized(ids=[], vocab=qvocab))
    assert len(flat) == 0 and ptr.tolist() == [0]


def test_int_query_filter_equals_reference_loop():
    """The table-based filter of list[list[int]] queries equals the reference's per-token set filter + empty-token rule
    (bm25s/__init__.py:778-798)."""
    from bm25s_b200.bm25 import _flatten

    r, g = _retriever_with_scores()
    rng = np.random.default_rng(4)
    r.vocab_dict = {f"w{i}": i for i in range(0, 50, 2)}
    r.vocab_dict[""] = 51
    r.unique_token_ids_set = set(r.vocab_dict.values())
    queries = [rng.integers(-2, 60, int(n)).tolist() for n in rng.integers(0, 9, 400)]
    flat, ptr = r._filter_int_queries(queries)
    ref = []
    for q in queries:
        qf = [t for t in q if t in r.unique_token_ids_set]
        ref.append(qf if qf else [r.vocab_dict[""]])
    ref_flat, ref_ptr = _flatten(ref)
    np.testing.assert_array_equal(flat, ref_flat)
    np.testing.assert_array_equal(ptr, ref_ptr)
    assert any(len([t for t in q if t in r.unique_token_ids_set]) == 0 for q in queries)
    del r.vocab_dict[""]
    r.unique_token_ids_set = set(r.vocab_dict.values())
    with pytest.raises(ValueError, match="does not contain any tokens that are in the vocabulary"):
        r._filter_int_queries([[1, 3]])
    flat, ptr = r._filter_int_queries([[0, 1, 2, 2]])
    assert flat.tolist() == [0, 2, 2] and ptr.tolist() == [0, 3]
