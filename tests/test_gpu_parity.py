"""GPU parity tests proper: the CUDA path (through the C ABI / the BM25 class) against the golden fixtures made
from the reference and against the oracles on seeded inputs. Integer/index results bit-exact (ids as sets within
runs of equal score); float32 scores bit-exact (tolerance 0) on the query path."""
import os

import numpy as np
import pytest

from _check import check_topk_rows, lists_from_flat, load_golden
from oracle import c_oracle as CO
from oracle import oracle as O

pytestmark = pytest.mark.gpu

METHODS = O.METHODS


def key(m):
    return m.replace("+", "plus")


def _dev(g, m, **kw):
    from bm25s_b200 import DeviceIndex

    nonocc = g[f"{key(m)}_nonocc"] if f"{key(m)}_nonocc" in g else None
    N = len(g["corpus_ptr"]) - 1
    return DeviceIndex.upload(g[f"{key(m)}_data"], g[f"{key(m)}_indices"], g[f"{key(m)}_indptr"], N, nonocc=nonocc, **kw), nonocc, N


def _shifts(nonocc, queries):
    return None if nonocc is None else np.array([O.query_shift(nonocc, q) for q in queries], np.float32)


@pytest.mark.parametrize("name", ["synth_small", "synth_medium"])
@pytest.mark.parametrize("force_general", [0, 1])
def test_retrieve_matches_reference_fixture(name, force_general, monkeypatch):
    monkeypatch.setenv("BM25CU_FORCE_GENERAL", str(force_general))
    g = load_golden(name)
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    k = int(g["k"])
    for m in METHODS:
        dev, nonocc, N = _dev(g, m)
        data, indices, indptr = g[f"{key(m)}_data"], g[f"{key(m)}_indices"], g[f"{key(m)}_indptr"]
        sh = _shifts(nonocc, queries)
        ids, sc, st = dev.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh, with_stats=True)
        dense_fn = lambda qi: O.get_scores(data, indptr, indices, N, queries[qi], nonocc)  # noqa: E731
        check_topk_rows(ids, sc, g[f"{key(m)}_ids"], g[f"{key(m)}_scores"], dense_fn, what=f"{name} {m} general={force_general}")
        assert st["n_general"] == (len(queries) if force_general else 0)
        assert st["posting_bytes"] == 8 * sum(int(indptr[t + 1] - indptr[t]) for q in queries for t in q)
        # weight masks (numpy order: mask, then shift): 0/1 float32 and arbitrary float64
        for mk in ("mask01", "maskf64"):
            ids, sc = dev.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh, mask=g[mk])
            np.testing.assert_array_equal(sc, g[f"{key(m)}_scores_{mk}"])
            dfn = lambda qi: O.get_scores(data, indptr, indices, N, queries[qi], nonocc, g[mk])  # noqa: E731
            check_topk_rows(ids, sc, ids, g[f"{key(m)}_scores_{mk}"], dfn, what=f"{name} {m} {mk}")
        dev.close()


@pytest.mark.parametrize("name", ["readme_corpus", "synth_small"])
def test_dense_scores_bit_exact(name):
    g = load_golden(name)
    for m in METHODS:
        dev, nonocc, N = _dev(g, m)
        if name == "readme_corpus":
            q = g["query_ids"]
            sh = None if nonocc is None else O.query_shift(nonocc, q)
            np.testing.assert_array_equal(dev.scores_dense(q, shift=sh), g[f"{key(m)}_dense"])
            for mk, mask in (("f32", np.array([1, 0, 0, 1], np.float32)), ("i32", np.array([1, 0, 0, 1], np.int32)),
                             ("bool", np.array([1, 0, 0, 1], np.bool_))):
                np.testing.assert_array_equal(dev.scores_dense(q, shift=sh, mask=mask.astype(np.float64)),
                                              g[f"{key(m)}_dense_mask_{mk}"])
        else:
            queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
            for qi in range(g[f"{key(m)}_dense"].shape[0]):
                sh = None if nonocc is None else O.query_shift(nonocc, queries[qi])
                np.testing.assert_array_equal(dev.scores_dense(queries[qi], shift=sh), g[f"{key(m)}_dense"][qi])
                np.testing.assert_array_equal(dev.scores_dense(queries[qi], shift=sh, mask=g["maskf64"]),
                                              g[f"{key(m)}_dense_maskf64"][qi])
        dev.close()


def test_known_answer_csc():
    """tests/core/test_scoring.py:26-98 through the ABI."""
    from bm25s_b200 import DeviceIndex

    dev = DeviceIndex.upload(np.array([1.0, 2.0], np.float32), np.array([0, 1], np.int32), np.array([0, 1, 2], np.int64), 3)
    np.testing.assert_array_equal(dev.scores_dense([0, 1]), [1.0, 2.0, 0.0])
    np.testing.assert_array_equal(dev.scores_dense([]), [0.0, 0.0, 0.0])
    ids, sc = dev.retrieve([0, 1], [0, 2], 3)
    np.testing.assert_array_equal(ids, [[1, 0, 2]])
    np.testing.assert_array_equal(sc, [[2.0, 1.0, 0.0]])
    dev2 = DeviceIndex.upload(np.array([1.5, 2.5], np.float32), np.array([0, 0], np.int32), np.array([0, 1, 2], np.int64), 2)
    np.testing.assert_array_equal(dev2.scores_dense([0, 1]), [4.0, 0.0])
    ids, sc = dev2.retrieve([0, 1], [0, 2], 2)
    np.testing.assert_array_equal(ids, [[0, 1]])
    np.testing.assert_array_equal(sc, [[4.0, 0.0]])
    with pytest.raises(ValueError, match="maximum token ID"):
        dev2.retrieve([0, 2], [0, 2], 1)


def test_topk_dense_reference_vectors():
    """tests/core/test_topk.py:16-59 and tests/core/test_selection.py:34-58 on the GPU selector."""
    from bm25s_b200 import selection

    g = load_golden("known_answers")
    s, i = selection.topk(g["topk_scores_in"], 5)
    np.testing.assert_array_equal(s, g["topk_s"])
    np.testing.assert_array_equal(i, g["topk_i"])
    assert i.dtype == np.int64 and s.dtype == np.float64
    s, i = selection.topk(np.array([3.0, 1.0, 4.0, 2.0], np.float32), k=2)
    np.testing.assert_array_equal(s, [4.0, 3.0])
    np.testing.assert_array_equal(i, [2, 0])
    rng = np.random.default_rng(3)
    for n, k in ((1, 1), (33, 33), (5000, 100), (200_000, 1000), (70_000, 5000)):
        x = rng.standard_normal(n).astype(np.float32)
        x[rng.integers(0, n, n // 3)] = 0.0  # many ties at a common value
        s, i = selection.topk(x, k)
        ref = np.sort(x)[-k:][::-1]
        np.testing.assert_array_equal(s, ref)
        np.testing.assert_array_equal(x[i], s)
        assert len(set(i.tolist())) == k


def _random_csc(rng, n_docs, n_vocab, nnz_per_col_max, with_zero=False):
    cols = []
    for t in range(n_vocab):
        df = int(rng.integers(0, min(nnz_per_col_max, n_docs) + 1))
        cols.append(np.sort(rng.choice(n_docs, size=df, replace=False)).astype(np.int32))
    indptr = np.zeros(n_vocab + 1, np.int64)
    np.cumsum([len(c) for c in cols], out=indptr[1:])
    indices = np.concatenate(cols) if indptr[-1] else np.zeros(0, np.int32)
    data = (rng.random(len(indices)) * 3).astype(np.float32)
    if with_zero and len(data):
        data[rng.integers(0, len(data), len(data) // 10)] = 0.0
    return data, indices.astype(np.int32), indptr


def test_edge_cases_small():
    """Empty queries, k = 0..N, repeated tokens, > 64 tokens (general path), all-zero columns, fewer than k matches."""
    from bm25s_b200 import DeviceIndex

    rng = np.random.default_rng(5)
    n_docs, n_vocab = 37, 90
    data, indices, indptr = _random_csc(rng, n_docs, n_vocab, 20, with_zero=True)
    dev = DeviceIndex.upload(data, indices, indptr, n_docs)
    queries = [[], [3], [3, 3, 3], [5, 7, 5, 9, 5], list(range(80)), [int(x) for x in rng.integers(0, n_vocab, 70)],
               [11, 12], [], [int(x) for x in rng.integers(0, n_vocab, 15)], [int(x) for x in rng.integers(0, n_vocab, 16)]]
    ptr = np.zeros(len(queries) + 1, np.int64)
    np.cumsum([len(q) for q in queries], out=ptr[1:])
    flat = np.array([t for q in queries for t in q], np.int32)
    for k in (0, 1, 2, 5, 36, 37):
        ids, sc, st = dev.retrieve(flat, ptr, k, with_stats=True)
        assert ids.shape == (len(queries), k)
        if k == 0:
            continue
        assert st["n_general"] == sum(len(q) > 15 for q in queries)  # queries longer than 15 tokens take the general path
        rid, rsc = O.retrieve(data, indptr, indices, n_docs, queries, k)
        dense_fn = lambda qi: O.score_dense(data, indptr, indices, n_docs, queries[qi])  # noqa: E731
        check_topk_rows(ids, sc, rid, rsc, dense_fn, what=f"edge k={k}")
    # a shard with fewer documents than k pads with id -1 / -inf
    dev.close()


@pytest.mark.parametrize("k", [1, 10, 100, 1000])
def test_random_corpus_vs_c_oracle(k):
    """Zipf corpus with heavy columns: windows, contested documents, candidate-buffer compaction, k up to 1000."""
    from bm25s_b200 import DeviceIndex, synth

    n_docs, n_vocab = 120_000, 20_000
    corp = synth.make_corpus(n_docs, 30, n_vocab, seed=7)
    data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "bm25+")
    qs = synth.make_queries(64, n_vocab, seed=8)
    # make some queries heavy: the most frequent tokens, repeated tokens
    extra = [[0, 1, 2], [0, 0], [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11], [5]]
    queries = qs.to_lists() + extra
    ptr = np.zeros(len(queries) + 1, np.int64)
    np.cumsum([len(q) for q in queries], out=ptr[1:])
    flat = np.array([t for q in queries for t in q], np.int32)
    sh = np.array([O.query_shift(nonocc, q) for q in queries], np.float32)
    dev = DeviceIndex.upload(data, indices, indptr, n_docs, nonocc=nonocc)
    ids, sc, st = dev.retrieve(flat, ptr, k, shift=sh, with_stats=True)
    assert st["n_general"] == 0
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, flat, ptr, k, shift=sh)
    dense_fn = lambda qi: CO.scores_dense(data, indices, indptr, n_docs, queries[qi], shift=sh[qi])  # noqa: E731
    check_topk_rows(ids, sc, rid, rsc, dense_fn if k <= 100 else None, what=f"random k={k}")
    # small stage / tag settings force many windows and the same answer
    # BM25CU_MS=0: the streaming kernel alone (retrieve_fast.cu); the other sets steer the list-at-a-time kernel
    # (retrieve_ms.cu, k <= 32): one part per query, many small parts, a tiny survivor budget (queries handed on to the
    # streaming kernel), lookups without bucket tables / presence bitmaps
    for env in ({"BM25CU_MS": "0"}, {"BM25CU_MS": "0", "BM25CU_POOL": "1024", "BM25CU_TAG": "4096"},
                {"BM25CU_MS": "0", "BM25CU_THREADS": "256", "BM25CU_CTAS_PER_SM": "2"},
                {"BM25CU_MS": "0", "BM25CU_MAXSHIFT": "0"}, {"BM25CU_MS": "0", "BM25CU_PRUNE": "0"},
                {"BM25CU_MS": "0", "BM25CU_MAXSHIFT": "1", "BM25CU_TAG": "8192"},
                {"BM25CU_MS": "0", "BM25CU_THREADS": "128", "BM25CU_PRUNE": "0"}, {"BM25CU_MS": "0", "BM25CU_THREADS": "1024"},
                {"BM25CU_MS_MAXPARTS": "1"}, {"BM25CU_MS_PARTCOST": "512"}, {"BM25CU_MS_PARTCOST": "2000", "BM25CU_MS_MAXPARTS": "255"},
                {"BM25CU_MS_BUDGET": "64"}, {"BM25CU_MS_BUDGET": "600", "BM25CU_MS_PARTCOST": "4096"},
                {"BM25CU_MS_BTAB": "0"}, {"BM25CU_MS_PBM": "0"}, {"BM25CU_MS_BTAB": "0", "BM25CU_MS_PBM": "0", "BM25CU_SEED_KTH": "0"},
                # lookups without the slot tables (bucket table + id slice + score), slot tables with many parts
                {"MS_STAB": 0}, {"MS_STAB": 0, "MS_PBM": 0}, {"MS_PARTCOST": 512, "MS_PBM": 0},
                # top lists in shared memory (k > 32): 128-thread CTAs, early / late merges of the pending keys
                {"MS_BIG_NT": 128}, {"MS_BIG_NT": 128, "MS_BIGMERGE": 1}, {"MS_BIGMERGE": 64, "MS_PARTCOST": 2048}):
        with dev.options(**env):
            ids2, sc2 = dev.retrieve(flat, ptr, k, shift=sh)
        np.testing.assert_array_equal(sc2, sc)
        np.testing.assert_array_equal(ids2, ids)  # deterministic tie rule: identical ids whatever the tiling
    dev.close()
    # slot tables in the compact sector format (16-bit doc offsets, five slots per sector; chosen when the handle is made),
    # also at a load that makes sectors lend slots and overflow to the column
    for env in ({"BM25CU_STAB_COMPACT": "1"}, {"BM25CU_STAB_COMPACT": "1", "BM25CU_STAB_LOAD": "80"}, {"BM25CU_STAB_LOAD": "80"}):
        old = {kk: os.environ.get(kk) for kk in env}
        os.environ.update(env)
        try:
            dev_c = DeviceIndex.upload(data, indices, indptr, n_docs, nonocc=nonocc)
        finally:
            for kk, v in old.items():
                if v is None:
                    os.environ.pop(kk, None)
                else:
                    os.environ[kk] = v
        for opts in ({}, {"MS_PARTCOST": 512}, {"MS_PBM": 0}):
            with dev_c.options(**opts):
                ids_c, sc_c = dev_c.retrieve(flat, ptr, k, shift=sh)
            np.testing.assert_array_equal(sc_c, sc, err_msg=str((env, opts)))
            np.testing.assert_array_equal(ids_c, ids, err_msg=str((env, opts)))
        dev_c.close()


def test_views_share_the_index_and_overlap():
    """bm25cu_index_view: handles on the same resident shard with streams and workspaces of their own - calls enqueued on
    the handle and on two views at the same time return the rows a single call returns (k in registers and in shared
    memory, and the streaming kernel); freeing the views leaves the parent usable."""
    import torch

    from bm25s_b200 import DeviceIndex, synth

    n_docs, n_vocab = 200_000, 20_000
    corp = synth.make_corpus(n_docs, 30, n_vocab, seed=17)
    data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "lucene")
    dev = DeviceIndex.upload(data, indices, indptr, n_docs)
    views = [dev.view(), dev.view()]
    lanes = [dev] + views
    dt = torch.device("cuda", 0)
    batches = []
    for s in range(3):
        qs = synth.make_queries(3000, n_vocab, seed=30 + s)
        batches.append((torch.from_numpy(qs.tokens).to(dt), torch.from_numpy(qs.ptr).to(dt), qs))
    torch.cuda.synchronize()
    for k in (10, 100):
        for opts in ({}, {"MS": 0}):
            want = []
            for qt, qp, qs in batches:
                with dev.options(**opts):
                    want.append(dev.retrieve(qs.tokens, qs.ptr, k))
            outs = [(torch.empty((3000, k), dtype=torch.int32, device=dt), torch.empty((3000, k), dtype=torch.float32, device=dt)) for _ in lanes]
            for ln in lanes:
                for kk, v in opts.items():
                    ln.set_option(kk, v)
            for rep in range(3):  # all three lanes in flight, three rounds, every lane sees every batch
                for li, ln in enumerate(lanes):
                    qt, qp, qs = batches[(li + rep) % 3]
                    ln.retrieve_device_enqueue(qt.data_ptr(), qp.data_ptr(), 3000, int(qt.numel()), k, outs[li][0].data_ptr(), outs[li][1].data_ptr())
                for li, ln in enumerate(lanes):
                    qt, qp, qs = batches[(li + rep) % 3]
                    ln.collect(3000, int(qt.numel()), k)
                    np.testing.assert_array_equal(outs[li][1].cpu().numpy(), want[(li + rep) % 3][1], err_msg=f"k={k} {opts} lane {li}")
                    np.testing.assert_array_equal(outs[li][0].cpu().numpy(), want[(li + rep) % 3][0], err_msg=f"k={k} {opts} lane {li}")
            for ln in lanes:
                for kk in opts:
                    ln.unset_option(kk)
    for v in views:
        v.close()
    ids, sc = dev.retrieve(batches[0][2].tokens, batches[0][2].ptr, 10)
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, batches[0][2].tokens, batches[0][2].ptr, 10)
    check_topk_rows(ids, sc, rid, rsc, None, what="parent after its views were freed")
    dev.close()


def test_weight_mask_fast_path():
    """masks with values in [0, 1] stay on the fused kernel (applied where a candidate is emitted, numpy
    order: mask before shift); results equal the C oracle's and the dense general kernel's. Other masks are routed."""
    from bm25s_b200 import DeviceIndex, synth

    n_docs, n_vocab = 120_000, 20_000
    corp = synth.make_corpus(n_docs, 30, n_vocab, seed=21)
    data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "bm25l")
    queries = synth.make_queries(96, n_vocab, seed=22).to_lists() + [[0, 1, 2], [3, 3], list(range(10))]
    ptr = np.zeros(len(queries) + 1, np.int64)
    np.cumsum([len(q) for q in queries], out=ptr[1:])
    flat = np.array([t for q in queries for t in q], np.int32)
    sh = np.array([O.query_shift(nonocc, q) for q in queries], np.float32)
    rng = np.random.default_rng(23)
    dev = DeviceIndex.upload(data, indices, indptr, n_docs, nonocc=nonocc)
    masks = {"01": (rng.random(n_docs) < 0.3).astype(np.float64), "unit": rng.random(n_docs),
             "sparse": (rng.random(n_docs) < 0.0005).astype(np.float64), "zeros": np.zeros(n_docs)}
    for name, mask in masks.items():
        for k in (1, 10, 300):
            ids, sc, st = dev.retrieve(flat, ptr, k, shift=sh, mask=mask, with_stats=True)
            assert st["n_general"] == 0, name
            rid, rsc = CO.retrieve(data, indices, indptr, n_docs, flat, ptr, k, shift=sh, mask=mask)
            dense_fn = lambda qi: CO.scores_dense(data, indices, indptr, n_docs, queries[qi], shift=sh[qi], mask=mask)  # noqa: E731
            check_topk_rows(ids, sc, rid, rsc, dense_fn if k <= 10 else None, what=f"mask {name} k={k}")
            with dev.options(MASK_FAST=0):
                ids_g, sc_g, st_g = dev.retrieve(flat, ptr, k, shift=sh, mask=mask, with_stats=True)
            assert st_g["n_general"] == len(queries)
            np.testing.assert_array_equal(sc, sc_g, err_msg=f"mask {name} k={k}")
    for bad in (np.where(rng.random(n_docs) < 0.5, 2.0, 0.5), np.where(rng.random(n_docs) < 0.01, -1.0, 1.0)):
        ids, sc, st = dev.retrieve(flat, ptr, 10, shift=sh, mask=bad, with_stats=True)
        assert st["n_general"] == len(queries)  # values outside [0, 1]: the dense kernel
        rid, rsc = CO.retrieve(data, indices, indptr, n_docs, flat, ptr, 10, shift=sh, mask=bad)
        np.testing.assert_array_equal(sc, rsc)
    dev.close()


def test_resident_weight_masks_bitmap_and_float():
    """bm25cu_index_set_mask / bm25cu_group_set_mask: a 0/1 mask kept on the device as one bit per document
    (its documents are dropped before any lookup) and a float64 mask kept resident return exactly what the per-call
    float64 mask returns - the reference's own masked fixtures - on one device and sharded at unaligned boundaries; the
    mask stays until cleared."""
    from bm25s_b200 import DeviceGroup, DeviceIndex, _lib

    g = load_golden("synth_medium")
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    k = int(g["k"])
    N = len(g["corpus_ptr"]) - 1
    for m in ("lucene", "bm25+"):
        data, indices, indptr = g[f"{key(m)}_data"], g[f"{key(m)}_indices"], g[f"{key(m)}_indptr"]
        nonocc = g[f"{key(m)}_nonocc"] if f"{key(m)}_nonocc" in g else None
        sh = _shifts(nonocc, queries)
        handles = [DeviceIndex.upload(data, indices, indptr, N, nonocc=nonocc),
                   DeviceGroup.upload(data, indices, indptr, N, nonocc=nonocc, devices=[0, 0, 0])]
        for dev in handles:
            ids0, sc0 = dev.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh)
            for mk, want_kind in (("mask01", _lib.MASK_BITS), ("maskf64", _lib.MASK_F64)):
                kind, arr = _lib.pack_weight_mask(g[mk])
                assert kind == want_kind
                dev.set_mask(kind, arr)
                for variant in ({}, {"MS": 0}, {"FORCE_GENERAL": 1}, {"MS_PARTCOST": 512}):
                    with dev.options(**variant):
                        ids, sc, st = dev.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh, with_stats=True)
                    np.testing.assert_array_equal(sc, g[f"{key(m)}_scores_{mk}"], err_msg=f"{m} {mk} {variant}")
                    dfn = lambda qi: O.get_scores(data, indptr, indices, N, queries[qi], nonocc, g[mk])  # noqa: E731
                    check_topk_rows(ids, sc, ids, g[f"{key(m)}_scores_{mk}"], dfn, what=f"resident {m} {mk} {variant}")
                # a mask given with the call wins over the resident one
                ids_c, sc_c = dev.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh, mask=np.ones(N))
                np.testing.assert_array_equal(sc_c, sc0)
            dev.set_mask(_lib.MASK_NONE)
            ids1, sc1 = dev.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh)
            np.testing.assert_array_equal(sc1, sc0)
            np.testing.assert_array_equal(ids1, ids0)
            dev.close()


def test_bm25_weight_mask_dtypes_and_sticky_mask():
    """BM25.retrieve(weight_mask=...) with bool / int32 / float32 / float64 0-1 masks (bit-packed on the host, resident on
    the device) and a fractional mask equals the reference fixtures; a mask does not outlive its call unless it was set
    with set_weight_mask; a read-only mask object is uploaded once."""
    from bm25s_b200 import BM25, _lib

    g = load_golden("synth_small")
    corpus = lists_from_flat(g["corpus_tokens"], g["corpus_ptr"])
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    V, k = int(g["n_vocab"]), int(g["k"])
    r = BM25(method="bm25l", backend="cuda")
    r.index((corpus, {i: i for i in range(V)}), create_empty_token=False, show_progress=False)
    plain = r.retrieve(queries, k=k).scores
    want01, wantf = g["bm25l_scores_mask01"], g["bm25l_scores_maskf64"]
    for dt in (bool, np.int32, np.float32, np.float64):
        res = r.retrieve(queries, k=k, weight_mask=g["mask01"].astype(dt))
        np.testing.assert_array_equal(res.scores, want01, err_msg=str(dt))
        np.testing.assert_array_equal(r.retrieve(queries, k=k).scores, plain)  # gone with the call
    np.testing.assert_array_equal(r.retrieve(queries, k=k, weight_mask=g["maskf64"]).scores, wantf)
    ro = g["mask01"].astype(bool)
    ro.setflags(write=False)
    calls = []
    orig = _lib.DeviceIndex.set_mask
    _lib.DeviceIndex.set_mask = lambda self, kind, arr=None: (calls.append(kind), orig(self, kind, arr))[1]
    try:
        for _ in range(3):
            np.testing.assert_array_equal(r.retrieve(queries, k=k, weight_mask=ro).scores, want01)
        assert calls == [_lib.MASK_BITS]  # packed and uploaded once
    finally:
        _lib.DeviceIndex.set_mask = orig
    r.set_weight_mask(g["mask01"])
    np.testing.assert_array_equal(r.retrieve(queries, k=k).scores, want01)
    np.testing.assert_array_equal(r.retrieve(queries, k=k).scores, want01)
    r.clear_weight_mask()
    np.testing.assert_array_equal(r.retrieve(queries, k=k).scores, plain)
    with pytest.raises(ValueError):
        r.retrieve(queries, k=k, weight_mask=np.ones(3))


def test_bm25_retrieve_from_several_threads():
    """BM25.retrieve called from several threads at once (a server): unmasked calls overlap on handles of their own (the
    index's handle + views, at most MAX_CALLS_IN_FLIGHT), masked calls take turns on the handle the mask lives on - every
    call returns what it returns alone, whatever runs next to it."""
    import threading

    from bm25s_b200 import BM25

    g = load_golden("synth_medium")
    corpus = lists_from_flat(g["corpus_tokens"], g["corpus_ptr"])
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    V, k = int(g["n_vocab"]), int(g["k"])
    r = BM25(method="bm25+", backend="cuda")
    r.index((corpus, {i: i for i in range(V)}), create_empty_token=False, show_progress=False)
    halves = [queries[: len(queries) // 2], queries[len(queries) // 2:], queries]
    want_plain = [r.retrieve(q, k=k) for q in halves]
    want_mask = [r.retrieve(q, k=k, weight_mask=g["mask01"]) for q in halves]
    want_maskf = [r.retrieve(q, k=k, weight_mask=g["maskf64"]) for q in halves]
    np.testing.assert_array_equal(want_mask[2].scores, g["bm25plus_scores_mask01"])
    np.testing.assert_array_equal(want_plain[2].scores, g["bm25plus_scores"])
    failures = []

    def caller(t):
        try:
            for it in range(12):
                j = (t + it) % 3
                kind = (t + it // 3) % 4 if t >= 4 else 0  # threads 0-3: unmasked only; 4-5: a mix
                if kind in (0, 3):
                    got, want = r.retrieve(halves[j], k=k), want_plain[j]
                elif kind == 1:
                    got, want = r.retrieve(halves[j], k=k, weight_mask=g["mask01"]), want_mask[j]
                else:
                    got, want = r.retrieve(halves[j], k=k, weight_mask=g["maskf64"]), want_maskf[j]
                np.testing.assert_array_equal(got.scores, want.scores, err_msg=f"thread {t} call {it} kind {kind}")
                np.testing.assert_array_equal(got.documents, want.documents, err_msg=f"thread {t} call {it} kind {kind}")
        except BaseException as e:  # noqa: BLE001
            failures.append(e)

    pool = [threading.Thread(target=caller, args=(t,)) for t in range(6)]
    for th in pool:
        th.start()
    for th in pool:
        th.join()
    if failures:
        raise failures[0]
    assert r._lanes is not None and len(r._lanes["views"]) <= BM25.MAX_CALLS_IN_FLIGHT - 1
    r.invalidate()  # frees the views, then the handle; the next call uploads again
    np.testing.assert_array_equal(r.retrieve(queries, k=k).scores, want_plain[2].scores)


def test_checked_build_finds_no_out_of_range_access():
    """`make checked` (bm25s_b200/libbm25cu_checked.so: -DBM25CU_CHECKED, device-side bounds checks of every index the
    list-at-a-time kernel computes; compute-sanitizer is not available on the GPU pool): the workload of
    tools/sanitizer_workload.py - every kernel path, masks, parts, several shards - runs clean and equals the oracle."""
    import subprocess
    import sys

    from bm25s_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(os.path.dirname(_lib.SO_PATH), "libbm25cu_checked.so")
    if not os.path.exists(so):
        pytest.skip("libbm25cu_checked.so has not been built (make -C bm25s_b200/csrc checked)")
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitizer_workload.py")], capture_output=True, text=True,
                         timeout=900, cwd=root, env=dict(os.environ, BM25CU_LIBRARY="checked"))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "sanitizer workload ok" in res.stdout and "libbm25cu_checked.so" in res.stdout


def test_upload_validation_and_bad_token_ids_on_device_entry_points():
    """(i) an uploaded CSC is validated on the device - non-monotone indptr, ids outside [0, n_docs), rows
    that do not ascend are refused with ValueError instead of being searched; (ii) the device entry points (no host-side
    check of the tokens) report an out-of-range token id for EVERY route - list-at-a-time kernel, streaming kernel, and
    the dense kernel that serves queries of more than 15 tokens - and never index indptr with it; (iii) a k the
    list-at-a-time kernel cannot take skips it altogether (no partial-result buffer) and still returns the oracle's rows."""
    import torch

    from bm25s_b200 import DeviceGroup, DeviceIndex, synth

    n_docs, n_vocab = 30_000, 5_000
    corp = synth.make_corpus(n_docs, 30, n_vocab, seed=31)
    data, indices, indptr, _ = O.build_index(corp.to_lists(), n_vocab, "lucene")
    # (i)
    bad_ptr = indptr.copy()
    bad_ptr[10], bad_ptr[11] = bad_ptr[11], bad_ptr[10] - 1 if bad_ptr[10] > 0 else 0
    t = int(np.argmax(np.diff(indptr) >= 4))
    swapped = indices.copy()
    swapped[indptr[t]], swapped[indptr[t] + 1] = swapped[indptr[t] + 1], swapped[indptr[t]]
    out_of_range = indices.copy()
    out_of_range[indptr[t]] = n_docs + 5
    for what, (d_, i_, p_) in {"indptr": (data, indices, bad_ptr), "unsorted": (data, swapped, indptr), "range": (data, out_of_range, indptr)}.items():
        with pytest.raises(ValueError):
            DeviceIndex.upload(d_, i_, p_, n_docs)
        if what != "indptr":  # (the host-side cut of the group upload walks the columns: a broken indptr is refused before it)
            with pytest.raises(ValueError):
                DeviceGroup.upload(d_, i_, p_, n_docs, devices=[0, 0])
    with pytest.raises(ValueError):
        DeviceGroup.upload(data, indices, bad_ptr, n_docs, devices=[0, 0])
    # (ii)
    dev = DeviceIndex.upload(data, indices, indptr, n_docs)
    dev_t = torch.device("cuda")
    good = [[1, 2, 3], [7], list(range(20, 40))]
    for bad_query in ([4, n_vocab + 3], [n_vocab], list(range(20, 39)) + [n_vocab + 1]):
        queries = good + [bad_query]
        flat = torch.tensor([t_ for q in queries for t_ in q], dtype=torch.int32, device=dev_t)
        ptr = torch.tensor(np.cumsum([0] + [len(q) for q in queries]), dtype=torch.int64, device=dev_t)
        oi = torch.empty((len(queries), 5), dtype=torch.int32, device=dev_t)
        osc = torch.empty((len(queries), 5), dtype=torch.float32, device=dev_t)
        for opts in ({}, {"MS": 0}, {"FORCE_GENERAL": 1}):
            with dev.options(**opts):
                with pytest.raises(ValueError, match="maximum token ID"):
                    dev.retrieve_device(flat.data_ptr(), ptr.data_ptr(), len(queries), int(flat.numel()), 5, oi.data_ptr(), osc.data_ptr())
    # the handle is still good
    qs = synth.make_queries(20, n_vocab, seed=32)
    ids, sc = dev.retrieve(qs.tokens, qs.ptr, 10)
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, qs.tokens, qs.ptr, 10)
    np.testing.assert_array_equal(sc, rsc)
    # (iii)
    ids, sc, st = dev.retrieve(qs.tokens, qs.ptr, 1500, with_stats=True)
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, qs.tokens, qs.ptr, 1500)
    np.testing.assert_array_equal(sc, rsc)
    assert st["ms_kernel_ms"] == 0  # k > 1024: the list-at-a-time kernel (plan, parts, partial rows) is not launched at all
    dev.close()


def test_kth_largest_table_and_seeded_threshold():
    """The upload-time table of per-column 10th / 100th / 1000th largest scores (start value of the pruning threshold)
    equals numpy's, and results with and without the seeded threshold are identical."""
    import ctypes as C

    from bm25s_b200 import DeviceIndex, synth, _lib

    n_docs, n_vocab = 60_000, 200_000
    corp = synth.make_corpus(n_docs, 40, n_vocab, seed=11)
    data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "lucene")
    dev = DeviceIndex.upload(data, indices, indptr, n_docs)
    out = np.zeros((3, n_vocab), np.float32)
    _lib.check(_lib.lib().bm25cu_debug_colkth(dev._h, out.ctypes.data_as(C.POINTER(C.c_float))))
    for ri, r in enumerate((10, 100, 1000)):
        ref = np.zeros(n_vocab, np.float32)
        for t in range(n_vocab):
            col = data[indptr[t]:indptr[t + 1]]
            if len(col) >= r:
                ref[t] = np.partition(col, len(col) - r)[len(col) - r]
        np.testing.assert_array_equal(out[ri], ref, err_msg=f"rank {r}")
    assert (out[2] > 0).any() and (out[0] == 0).any()  # both long and short columns were exercised
    queries = synth.make_queries(200, n_vocab, seed=12).to_lists() + [[0, 1, 2, 3], [7, 7]]
    ptr = np.zeros(len(queries) + 1, np.int64)
    np.cumsum([len(q) for q in queries], out=ptr[1:])
    flat = np.array([t for q in queries for t in q], np.int32)
    for k in (1, 5, 10, 50, 100, 700, 1000, 1500):
        ids, sc = dev.retrieve(flat, ptr, k)
        with dev.options(SEED_KTH=0):
            ids0, sc0 = dev.retrieve(flat, ptr, k)
        np.testing.assert_array_equal(sc, sc0, err_msg=f"k={k}")
        np.testing.assert_array_equal(ids, ids0, err_msg=f"k={k}")
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, flat, ptr, 10)
    ids, sc = dev.retrieve(flat, ptr, 10)
    np.testing.assert_array_equal(sc, rsc)
    dev.close()


def test_doc_range_shards_merge_to_unsharded():
    """shard by contiguous doc range, local top-k per shard, merge == unsharded result."""
    from bm25s_b200 import DeviceIndex, topk_merge

    g = load_golden("synth_medium")
    m = "bm25l"
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    nonocc = g[f"{key(m)}_nonocc"]
    N, k = len(g["corpus_ptr"]) - 1, int(g["k"])
    sh = _shifts(nonocc, queries)
    full, _, _ = _dev(g, m)
    ids_full, sc_full = full.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh)
    for parts in (2, 3, 8):
        bounds = [N * p // parts for p in range(parts + 1)]
        pi, ps = [], []
        for p in range(parts):
            d, _, _ = _dev(g, m, doc_lo=bounds[p], doc_hi=bounds[p + 1])
            assert d.n_docs == bounds[p + 1] - bounds[p] and d.doc_lo == bounds[p]
            i, s = d.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh)
            pi.append(i)
            ps.append(s)
            d.close()
        mi, ms = topk_merge(np.stack(pi), np.stack(ps))
        np.testing.assert_array_equal(ms, sc_full)
        np.testing.assert_array_equal(mi, ids_full)
    # shard download returns the shard-local CSC
    d, _, _ = _dev(g, m, doc_lo=100, doc_hi=1100)
    data, indices, indptr, no = d.download()
    full_idx = g[f"{key(m)}_indices"]
    sel = (full_idx >= 100) & (full_idx < 1100)
    np.testing.assert_array_equal(indices, full_idx[sel] - 100)
    np.testing.assert_array_equal(data, g[f"{key(m)}_data"][sel])
    np.testing.assert_array_equal(no, nonocc)
    full.close()


FULL_SIZE = [
    # BASELINE.json configs at their full sizes: (shape, method, k, batch)
    ("nq", "lucene", 10, 3_452),      # configs[1]
    ("nq", "bm25l", 10, 10_000),      # configs[3]: non-occurrence shift, 10k-query batch
    ("nq", "bm25+", 10, 10_000),      # configs[3]
    ("msmarco", "lucene", 10, 6_980),    # configs[2], the headline
    ("msmarco", "lucene", 1000, 6_980),  # configs[2], k = 1000
    ("msmarco", "lucene", 10, -6_980),   # the headline with COMPACT slot tables (16-bit doc offsets; a negative batch size marks it)
]


@pytest.mark.parametrize("shape,method,k,n_queries", FULL_SIZE)
def test_full_size_against_the_oracle(shape, method, k, n_queries, monkeypatch):
    """BASELINE.json's configurations at FULL size against the C oracle (the reference algorithm, oracle/bm25_oracle.c)
    on the same CSC: the index is built on the GPU, downloaded, and a strided sample of >= 300 queries of the batch is
    scored by the oracle on the host cores (all of the matrix, dense vector + selection, as the reference does).
    Scores bit-equal, ids equal as sets above the k-th score, returned ids carry their oracle dense score and nothing
    outside a row beats its tail (sub-sample). Matches the cross-backend equality test of the reference,
    tests/numba/test_numba_backend_retrieve.py:68-144. On top of it, size-independent properties over the WHOLE batch:
    the list-at-a-time kernel and the streaming kernel (two independent algorithms) return identical rows; rows are
    sorted by (score desc, id asc) without duplicates."""
    import torch

    from bm25s_b200 import DeviceIndex, synth

    if n_queries < 0:  # creation-time knob: read from the environment when the handle is made
        monkeypatch.setenv("BM25CU_STAB_COMPACT", "1")
        n_queries = -n_queries
    n_docs, avg_len, n_vocab, _ = synth.SHAPES[shape]
    dev_t = torch.device("cuda")
    tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev_t)
    ix = DeviceIndex.build_device(tokens.data_ptr(), ptr.data_ptr(), n_docs, n_vocab, method, method, 1.5, 0.75, 0.5)
    assert ix.nnz > 0.9 * n_docs * avg_len  # the corpus really has the named shape (most tokens of a document are distinct)
    del tokens, ptr
    torch.cuda.empty_cache()
    qt, qp = synth.make_queries_torch(n_queries, n_vocab, seed=1, device=dev_t)
    qt_h, qp_h = qt.cpu().numpy(), qp.cpu().numpy()
    data, indices, indptr, nonocc = ix.download()
    sh_h = None if nonocc is None else np.array([O.query_shift(nonocc, qt_h[qp_h[i]:qp_h[i + 1]]) for i in range(n_queries)], np.float32)
    sh_d = None if sh_h is None else torch.from_numpy(sh_h).to(dev_t)

    def run(**opts):
        with ix.options(**opts):
            oi = torch.empty((n_queries, k), dtype=torch.int32, device=dev_t)
            osc = torch.empty((n_queries, k), dtype=torch.float32, device=dev_t)
            st = ix.retrieve_device(qt.data_ptr(), qp.data_ptr(), n_queries, int(qp_h[-1]), k, oi.data_ptr(), osc.data_ptr(),
                                    shift_ptr=None if sh_d is None else sh_d.data_ptr())
            return oi.cpu().numpy(), osc.cpu().numpy(), st

    ids1, sc1, st1 = run()
    assert st1["n_general"] == 0 and st1["ms_kernel_ms"] > 0
    # ---- the oracle on a strided sample of the batch
    step = max(1, n_queries // 320)
    sample = np.arange(0, n_queries, step)
    assert len(sample) >= 300
    s_ptr = np.zeros(len(sample) + 1, np.int64)
    np.cumsum([qp_h[q + 1] - qp_h[q] for q in sample], out=s_ptr[1:])
    s_tok = np.concatenate([qt_h[qp_h[q]:qp_h[q + 1]] for q in sample]).astype(np.int32)
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, s_tok, s_ptr, k, shift=None if sh_h is None else sh_h[sample])
    dense_every = max(1, len(sample) // 6)

    def dense_fn_for(j):
        q = sample[j]
        return CO.scores_dense(data, indices, indptr, n_docs, qt_h[qp_h[q]:qp_h[q + 1]], shift=None if sh_h is None else sh_h[q])

    check_topk_rows(ids1[sample], sc1[sample], rid, rsc, None, what=f"{shape} {method} k={k} vs oracle")
    for j in range(0, len(sample), dense_every):
        check_topk_rows(ids1[sample[j]:sample[j] + 1], sc1[sample[j]:sample[j] + 1], rid[j:j + 1], rsc[j:j + 1],
                        lambda _q, j=j: dense_fn_for(j), what=f"{shape} {method} k={k} dense q={sample[j]}")
    del data, indices, indptr
    # ---- whole batch: the streaming kernel alone returns the same rows
    if k <= 10:
        ids0, sc0, st0 = run(MS=0)
        assert st0["n_general"] == 0 and st0["ms_kernel_ms"] == 0  # the two runs really took different kernels
        np.testing.assert_array_equal(sc1.view(np.int32), sc0.view(np.int32))
        np.testing.assert_array_equal(ids1, ids0)
    assert np.all(sc1[:, :-1] >= sc1[:, 1:])
    tie = sc1[:, :-1] == sc1[:, 1:]
    assert np.all(ids1[:, :-1][tie] < ids1[:, 1:][tie])  # ties by ascending id
    ix.close()


@pytest.mark.parametrize("k", [10, 32, 40])
def test_peer_exchange_single_rank_epochs(k):
    """bm25cu_exchange_* with world = 1 (all a one-GPU box can run; bench.py checks N > 1 against all_gather + merge and
    the oracle): the push + merge kernels - one warp per query for k <= 32, one CTA per query above - return the rows in
    (score desc, id asc) order, call after call (both buffer parities), rows with fewer than k valid entries keep their
    -1 / -inf tail, and the BM25L / BM25+ shift is added after the merge (valid entries only)."""
    import torch

    from bm25s_b200 import _lib

    dev_t = torch.device("cuda")
    Q = 37
    g = torch.Generator(device="cpu").manual_seed(5)
    ex = _lib.PeerExchange(0, 0, 1, Q * k, lambda b: [b])
    for call in range(5):
        sc = torch.rand((Q, k), generator=g).round(decimals=1)  # ties on purpose
        ids = torch.stack([torch.randperm(1000, generator=g)[:k] for _ in range(Q)]).to(torch.int32)
        ids[3, 6:] = -1
        sc[3, 6:] = float("-inf")
        shift = torch.rand(Q, generator=g) if call % 2 else None
        d_ids, d_sc = ids.to(dev_t), sc.to(dev_t)
        d_shift = None if shift is None else shift.to(dev_t)
        o_ids, o_sc = torch.empty_like(d_ids), torch.empty_like(d_sc)
        ex.merge(d_ids.data_ptr(), d_sc.data_ptr(), Q, k, o_ids.data_ptr(), o_sc.data_ptr(), 0,
                 shift_ptr=None if d_shift is None else d_shift.data_ptr())
        ex.check(0)
        ref_i, ref_s = _lib.topk_merge(ids.numpy()[None], sc.numpy()[None], device=0)
        if shift is not None:
            ref_s = np.where(ref_i >= 0, (ref_s + shift.numpy()[:, None]).astype(np.float32), ref_s)
        np.testing.assert_array_equal(o_ids.cpu().numpy(), ref_i, err_msg=f"call {call}")
        np.testing.assert_array_equal(o_sc.cpu().numpy(), ref_s, err_msg=f"call {call}")
    ex.close()


def test_list_at_a_time_kernel_edge_cases():
    """retrieve_ms.cu against the C oracle on the shapes that steer its control flow: tokens whose column is empty, a
    query of one repeated token, 15 tokens (the most it takes), 16 (handed on to the general kernel), k on both sides of
    the register / shared-memory top list (32 | 33) and of its limit (1024 | 1025), k larger than the number of documents
    with a positive score (padding rule: handed on), every BM25 variant with its shift."""
    from bm25s_b200 import DeviceIndex, synth

    n_docs, n_vocab = 30_000, 3_000
    corp = synth.make_corpus(n_docs, 25, n_vocab, seed=31)
    docs = corp.to_lists()
    empty = sorted(set(range(n_vocab)) - set(t for d in docs for t in d))[:2]  # tokens without a posting, if any
    rare = [n_vocab - 1 - i for i in range(20)]
    queries = [[0], [5, 5, 5], [0, 1], [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14], list(range(16)),
               rare[:4], rare[:2] + [0, 1], [2, 700, 2, 1500, 2], [n_vocab - 1], [10, 20, 30, 40]]
    if empty:
        queries += [[empty[0]], [empty[0], 3], empty + [7, 7]]
    queries += synth.make_queries(40, n_vocab, seed=32).to_lists()
    ptr = np.zeros(len(queries) + 1, np.int64)
    np.cumsum([len(q) for q in queries], out=ptr[1:])
    flat = np.array([t for q in queries for t in q], np.int32)
    for method in ("lucene", "bm25l", "robertson"):
        data, indices, indptr, nonocc = O.build_index(docs, n_vocab, method)
        sh = None if nonocc is None else np.array([O.query_shift(nonocc, q) for q in queries], np.float32)
        dev = DeviceIndex.upload(data, indices, indptr, n_docs, nonocc=nonocc)
        for k in (1, 2, 31, 32, 33, 64, 1000, 1024, 1025, 5000):
            ids, sc, st = dev.retrieve(flat, ptr, k, shift=sh, with_stats=True)
            if k <= 2048 and data.min() >= 0:
                assert st["n_general"] == 1, (method, k)  # the 16-token query; larger k or negative scores: every query
            rid, rsc = CO.retrieve(data, indices, indptr, n_docs, flat, ptr, k, shift=sh)
            dense_fn = lambda qi: CO.scores_dense(data, indices, indptr, n_docs, queries[qi], shift=None if sh is None else sh[qi])  # noqa: E731
            check_topk_rows(ids, sc, rid, rsc, dense_fn if k <= 64 else None, what=f"{method} k={k}")
        dev.close()


@pytest.mark.parametrize("k", [1, 10, 50, 200])
def test_tie_heavy_corpus(k):
    """Every document has the same length and holds a token at most once, so a column carries ONE score value and whole
    runs of documents tie exactly - at the k-th place, too. The pruning rules keep bounds >= theta and accept keys > theta,
    so the rows must be the unique (score desc, id asc) top-k whatever the part layout; checked against the C oracle
    (scores bit-exact, ids as sets above the k-th score, every id verified against the dense scores) and for identical
    ids across kernels and part layouts."""
    from bm25s_b200 import DeviceIndex

    rng = np.random.default_rng(51)
    n_docs, n_vocab, L = 40_000, 600, 12
    # token t is drawn with probability ~ 1 / (t + 3): a few dense columns, many sparse ones
    p = 1.0 / (np.arange(n_vocab) + 3.0)
    p /= p.sum()
    docs = [rng.choice(n_vocab, size=L, replace=False, p=p).tolist() for _ in range(n_docs)]
    data, indices, indptr, nonocc = O.build_index(docs, n_vocab, "lucene")
    assert len(np.unique(data[indptr[0]:indptr[1]])) == 1  # one score per column: ties everywhere
    queries = [[0], [1, 2], [0, 1, 2], [5, 40, 300], [599], [2, 2, 7], [10, 11, 12, 13, 14, 15], [0, 599], [3, 250]]
    queries += [rng.choice(n_vocab, size=int(rng.integers(1, 6)), p=p).tolist() for _ in range(40)]
    ptr = np.zeros(len(queries) + 1, np.int64)
    np.cumsum([len(q) for q in queries], out=ptr[1:])
    flat = np.array([t for q in queries for t in q], np.int32)
    dev = DeviceIndex.upload(data, indices, indptr, n_docs)
    ids, sc = dev.retrieve(flat, ptr, k)
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, flat, ptr, k)
    dense_fn = lambda qi: CO.scores_dense(data, indices, indptr, n_docs, queries[qi])  # noqa: E731
    check_topk_rows(ids, sc, rid, rsc, dense_fn, what=f"ties k={k}")
    # the deterministic tie rule: within a run of equal scores the ids ascend, and nothing with a smaller id and the same
    # score as the row's tail was left out
    for q in range(len(queries)):
        tie = sc[q][:-1] == sc[q][1:]
        assert np.all(ids[q][:-1][tie] < ids[q][1:][tie]), q
        dense = dense_fn(q)
        tail = np.flatnonzero(dense == sc[q][-1])
        n_tail = int((sc[q] == sc[q][-1]).sum())
        assert set(ids[q][sc[q] == sc[q][-1]].tolist()) == set(tail[:n_tail].tolist()), q
    for env in ({"BM25CU_MS": "0"}, {"BM25CU_MS_MAXPARTS": "1"}, {"BM25CU_MS_PARTCOST": "256", "BM25CU_MS_MAXPARTS": "64"},
                {"BM25CU_SEED_KTH": "0"}, {"BM25CU_MS_PBM": "0", "BM25CU_MS_BTAB": "0"}, {"MS_STAB": 0}, {"MS_PBM": 0}):
        with dev.options(**env):
            ids2, sc2 = dev.retrieve(flat, ptr, k)
        np.testing.assert_array_equal(sc2, sc, err_msg=str(env))
        np.testing.assert_array_equal(ids2, ids, err_msg=str(env))
    dev.close()


def test_sharded_retrieve_with_peer_exchange_two_gpus():
    """Needs two GPUs (skipped on a one-GPU box): tools/sharded_check.py under torchrun - sharded GPU build, scoring,
    peer-memory exchange fused with the merge, and the NCCL variant, against the C oracle on the unsharded corpus."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29547", os.path.join(root, "tools", "sharded_check.py")],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "sharded check ok: bm25+ world 2" in res.stdout
