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


def test_weight_mask_packing():
    """0/1 masks of any dtype become one bit per document (little-endian bit order inside uint32 words: bit d % 32 of word
    d // 32, what bm25cu_index_set_mask(BM25CU_MASK_BITS) takes); anything else the exact float64 copy."""
    from bm25s_b200 import _lib

    rng = np.random.default_rng(0)
    keep = rng.random(1003) < 0.3
    for dt in (bool, np.int8, np.int32, np.int64, np.uint8, np.float32, np.float64):
        kind, words = _lib.pack_weight_mask(keep.astype(dt))
        assert kind == _lib.MASK_BITS and words.dtype == np.uint32 and len(words) >= (1003 + 31) // 32
        bits = (words[np.arange(1003) >> 5] >> (np.arange(1003) & 31)) & 1
        np.testing.assert_array_equal(bits.astype(bool), keep)
        assert not (words[1003 >> 5] >> np.uint32(1003 & 31)) & 1  # padding bits are clear
    for m in (np.where(keep, 0.5, 1.0), np.where(keep, 2, 0), np.where(keep, np.nan, 1.0), -keep.astype(np.float64)):
        kind, arr = _lib.pack_weight_mask(m)
        if np.array_equal(m, -keep.astype(np.float64)):  # -0.0 and -1.0: not a 0/1 mask unless nothing is kept
            assert kind == _lib.MASK_F64
        assert kind == _lib.MASK_F64 and arr.dtype == np.float64
        np.testing.assert_array_equal(arr, m.astype(np.float64))


def test_devices_argument_and_environment(monkeypatch):
    """BM25(devices=[...]) / BM25CU_DEVICES: the device list of the single-process multi-GPU form (parsed on the host; the
    shards are made at the first upload / build, which needs GPUs)."""
    from bm25s_b200 import BM25

    r = BM25(devices=[2, 3])
    assert r.devices == [2, 3] and r.device == 2
    monkeypatch.setenv("BM25CU_DEVICES", "0, 1,2")
    assert BM25().devices == [0, 1, 2]
    assert BM25(devices=[5]).devices == [5]  # the argument wins
    monkeypatch.delenv("BM25CU_DEVICES")
    assert BM25().devices is None
    with pytest.raises(ValueError):
        BM25(devices=[])
    # the residency key tells a sharded object from a single-device one
    r1, r2 = BM25(), BM25(devices=[0, 1])
    sc = {"data": np.zeros(1, np.float32), "indices": np.zeros(1, np.int32), "indptr": np.zeros(2, np.int64), "num_docs": 1}
    r1.scores = r2.scores = sc
    assert not BM25._same_key(r1._residency_key(), r2._residency_key())
    assert BM25._same_key(r1._residency_key(), r1._residency_key())


def test_bench_helpers_against_the_oracle():
    """bench.py's result digest, sub-batch extraction and oracle spot check (run on the oracle's own rows; a corrupted row
    must be reported)."""
    import importlib.util

    from bm25s_b200 import synth
    from oracle import c_oracle as CO
    from oracle import oracle as O

    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    n_docs, n_vocab, k = 20_000, 4_000, 10
    corp = synth.make_corpus(n_docs, 30, n_vocab, seed=5)
    data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "bm25l")
    qs = synth.make_queries(120, n_vocab, seed=6)
    shift = bench.query_shifts(nonocc, qs.tokens, qs.ptr)
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, qs.tokens, qs.ptr, k, shift=shift)
    tok, ptr = bench.sub_batch(qs.tokens, qs.ptr, [3, 7, 100])
    assert tok.tolist() == sum((qs.tokens[qs.ptr[q]:qs.ptr[q + 1]].tolist() for q in (3, 7, 100)), [])
    assert ptr.tolist() == np.cumsum([0] + [int(qs.ptr[q + 1] - qs.ptr[q]) for q in (3, 7, 100)]).tolist()
    ok = bench.oracle_parity((data, indices, indptr), qs.tokens, qs.ptr, n_docs, k, shift, rid.astype(np.int32), rsc, 60)
    assert ok["ok"] and ok["queries"] == 60
    bad = rsc.copy()
    bad[0, 0] = np.nextafter(bad[0, 0], np.float32(np.inf))
    rep = bench.oracle_parity((data, indices, indptr), qs.tokens, qs.ptr, n_docs, k, shift, rid.astype(np.int32), bad, 60)
    assert not rep["ok"] and rep["failures"]["score_rows"] >= 1 and rep["first_bad_query"]["q"] == 0
    assert bench.digest_of(rid, rsc) == bench.digest_of(rid.copy(), rsc.copy()) != bench.digest_of(rid, bad)


def test_handle_pool_of_concurrent_retrieve_calls():
    """BM25._lane (host logic, no GPU): one caller at a time always gets the index's own handle; concurrent callers get
    views, at most MAX_CALLS_IN_FLIGHT handles in all; a call that needs the handle the mask lives on waits for exactly
    that handle; handles of a device copy that was replaced are not handed out again."""
    import threading
    import time

    from bm25s_b200 import BM25, _lib

    class FakeIndex(_lib.DeviceIndex):
        made = 0

        def __init__(self):  # no library handle
            self.closed = False
            self.told = []

        def set_option(self, name, value):
            assert name == "IN_FLIGHT"
            self.told.append(value)

        def view(self):
            FakeIndex.made += 1
            return FakeIndex()

        def close(self):
            self.closed = True

    r = BM25(method="lucene")
    base = FakeIndex()
    r._dev = base
    for _ in range(3):
        with r._lane(base, base_only=False) as h:
            assert h is base
    assert FakeIndex.made == 0
    with r._lane(base, False) as h0, r._lane(base, False) as h1, r._lane(base, False) as h2:
        assert h0 is base and h1 is not base and h2 is not base and h1 is not h2
        assert base.told == [1] and h1.told == [2] and h2.told == [3]  # calls on the GPU when the handle was taken
        assert FakeIndex.made == 2 == BM25.MAX_CALLS_IN_FLIGHT - 1
        got, cms = [], []

        def fourth():
            cms.append(r._lane(base, False))
            got.append(cms[0].__enter__())

        th = threading.Thread(target=fourth)
        th.start()
        time.sleep(0.2)
        assert not got  # a fourth caller waits
    th.join(5)
    assert len(got) == 1 and FakeIndex.made == 2
    cms[0].__exit__(None, None, None)  # the fourth caller's handle goes back
    assert len(r._lanes["free"]) == 3 and r._lanes["free"][-1] is base
    # base_only waits for the base handle although views are free
    with r._lane(base, False) as h0:
        assert h0 is base
        got, cms = [], []

        def masked():
            cms.append(r._lane(base, True))
            got.append(cms[0].__enter__())

        th = threading.Thread(target=masked)
        th.start()
        time.sleep(0.2)
        assert not got
    th.join(5)
    assert got == [base]
    cms[0].__exit__(None, None, None)
    views = list(r._lanes["views"])
    r._close_dev()  # views first, then the handle
    assert all(v.closed for v in views) and base.closed and r._dev is None and r._lanes is None
    new = FakeIndex()
    with r._lane(new, False) as h:
        assert h is new and r._lanes["dev"] is new and r._lanes["views"] == []
