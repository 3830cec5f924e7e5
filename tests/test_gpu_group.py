"""One index over several devices from ONE process (include/bm25cu.h, bm25cu_group_*; BM25(devices=[...])): the rows must
be the single-device rows, whatever the number of shards - against the reference fixtures, the C oracle and the
single-device path. Device lists repeat device 0 when the box has one GPU (several shards on one GPU: the same code
path, peer copies become device-to-device copies); with two or more GPUs the real devices are used as well."""
import numpy as np
import pytest

from _check import check_topk_rows, lists_from_flat, load_golden
from oracle import c_oracle as CO
from oracle import oracle as O

pytestmark = pytest.mark.gpu

METHODS = O.METHODS


def key(m):
    return m.replace("+", "plus")


def device_lists():
    from bm25s_b200 import device_count

    n = device_count()
    lists = [[0, 0], [0, 0, 0], [0] * 8]
    if n >= 2:
        lists += [[0, 1], list(range(min(n, 8)))]
    return lists


@pytest.mark.parametrize("name", ["synth_small", "synth_medium"])
def test_group_upload_retrieve_matches_reference_fixture(name):
    from bm25s_b200 import DeviceGroup, DeviceIndex

    g = load_golden(name)
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    k = int(g["k"])
    N = len(g["corpus_ptr"]) - 1
    for m in METHODS:
        data, indices, indptr = g[f"{key(m)}_data"], g[f"{key(m)}_indices"], g[f"{key(m)}_indptr"]
        nonocc = g[f"{key(m)}_nonocc"] if f"{key(m)}_nonocc" in g else None
        sh = None if nonocc is None else np.array([O.query_shift(nonocc, q) for q in queries], np.float32)
        one = DeviceIndex.upload(data, indices, indptr, N, nonocc=nonocc)
        ids1, sc1 = one.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh)
        for devs in device_lists():
            grp = DeviceGroup.upload(data, indices, indptr, N, nonocc=nonocc, devices=devs)
            assert grp.n_shards == len(devs) and grp.nnz == len(data)
            ids, sc, st = grp.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh, with_stats=True)
            dense_fn = lambda qi: O.get_scores(data, indptr, indices, N, queries[qi], nonocc)  # noqa: E731
            check_topk_rows(ids, sc, g[f"{key(m)}_ids"], g[f"{key(m)}_scores"], dense_fn, what=f"{name} {m} group {devs}")
            # identical rows whatever the number of shards: the merge orders by the pre-shift scores, as one device does
            np.testing.assert_array_equal(sc, sc1)
            np.testing.assert_array_equal(ids, ids1)
            assert st["posting_bytes"] == 8 * sum(int(indptr[t + 1] - indptr[t]) for q in queries for t in q)
            for mk in ("mask01", "maskf64"):
                ids_m, sc_m = grp.retrieve(g["query_tokens"], g["query_ptr"], k, shift=sh, mask=g[mk])
                np.testing.assert_array_equal(sc_m, g[f"{key(m)}_scores_{mk}"], err_msg=f"{m} {mk} {devs}")
            # the whole matrix comes back in the reference's layout, dense scores cover every shard's range
            d2, i2, p2, n2 = grp.download()
            np.testing.assert_array_equal(d2, data)
            np.testing.assert_array_equal(i2, indices)
            np.testing.assert_array_equal(p2, indptr)
            if nonocc is not None:
                np.testing.assert_array_equal(n2, nonocc)
            np.testing.assert_array_equal(grp.scores_dense(np.array(queries[0], np.int32), shift=None if sh is None else sh[0]),
                                          g[f"{key(m)}_dense"][0])
            grp.close()
        one.close()


def test_group_build_is_bit_identical_to_the_reference():
    """bm25cu_group_build: shard-local counting, document frequencies summed over the shards, shard-local scoring - the
    downloaded matrix equals the reference's arrays bit for bit for every variant and shard count."""
    from bm25s_b200 import DeviceGroup

    g = load_golden("synth_medium")
    V = int(g["n_vocab"])
    for m in METHODS:
        for devs in device_lists()[:2] + device_lists()[3:]:
            grp = DeviceGroup.build(g["corpus_tokens"], g["corpus_ptr"], V, m, m, 1.5, 0.75, 0.5, devices=devs)
            d, i, p, n = grp.download()
            np.testing.assert_array_equal(p, g[f"{key(m)}_indptr"], err_msg=f"{m} {devs}")
            np.testing.assert_array_equal(i, g[f"{key(m)}_indices"], err_msg=f"{m} {devs}")
            np.testing.assert_array_equal(d, g[f"{key(m)}_data"], err_msg=f"{m} {devs}")
            if f"{key(m)}_nonocc" in g:
                np.testing.assert_array_equal(n, g[f"{key(m)}_nonocc"])
            grp.close()


@pytest.mark.parametrize("k", [10, 100, 1000])
def test_group_random_corpus_vs_c_oracle(k):
    from bm25s_b200 import DeviceGroup, synth

    n_docs, n_vocab = 120_000, 20_000
    corp = synth.make_corpus(n_docs, 30, n_vocab, seed=7)
    data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "bm25+")
    qs = synth.make_queries(64, n_vocab, seed=8)
    queries = qs.to_lists() + [[0, 1, 2], [0, 0], list(range(12)), [5], list(range(20))]
    ptr = np.zeros(len(queries) + 1, np.int64)
    np.cumsum([len(q) for q in queries], out=ptr[1:])
    flat = np.array([t for q in queries for t in q], np.int32)
    sh = np.array([O.query_shift(nonocc, q) for q in queries], np.float32)
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, flat, ptr, k, shift=sh)
    for devs in device_lists():
        grp = DeviceGroup.upload(data, indices, indptr, n_docs, nonocc=nonocc, devices=devs)
        ids, sc = grp.retrieve(flat, ptr, k, shift=sh)
        check_topk_rows(ids, sc, rid, rsc, None, what=f"group {devs} k={k}")
        with grp.options(MS=0):  # knobs reach every shard
            ids0, sc0 = grp.retrieve(flat, ptr, k, shift=sh)
        np.testing.assert_array_equal(sc0, sc)
        grp.close()
    with pytest.raises(ValueError):
        grp = DeviceGroup.upload(data, indices, indptr, n_docs, devices=[0, 0])
        grp.retrieve(np.array([n_vocab], np.int32), np.array([0, 1], np.int64), 5)


def test_bm25_class_over_several_devices(tmp_path, monkeypatch):
    """BM25(backend="cuda", devices=[...]): index / retrieve / get_scores / save / load with the index sharded over the
    device list - the arrays and rows of the single-device object."""
    from bm25s_b200 import BM25

    g = load_golden("synth_small")
    corpus = lists_from_flat(g["corpus_tokens"], g["corpus_ptr"])
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    V, k = int(g["n_vocab"]), int(g["k"])
    devs = device_lists()[-1]
    for m in ("lucene", "bm25l"):
        r = BM25(method=m, backend="cuda", devices=devs)
        r.index((corpus, {i: i for i in range(V)}), create_empty_token=False, show_progress=False)
        np.testing.assert_array_equal(r.scores["data"], g[f"{key(m)}_data"])
        np.testing.assert_array_equal(r.scores["indices"], g[f"{key(m)}_indices"])
        res = r.retrieve(queries, k=k, show_progress=False)
        check_topk_rows(res.documents, res.scores, g[f"{key(m)}_ids"], g[f"{key(m)}_scores"], None, what=f"BM25 devices {m}")
        np.testing.assert_array_equal(r.get_scores(queries[0]), g[f"{key(m)}_dense"][0])
        res_m = r.retrieve(queries, k=k, weight_mask=g["mask01"])
        np.testing.assert_array_equal(res_m.scores, g[f"{key(m)}_scores_mask01"])
        d = tmp_path / f"idx_{key(m)}"
        r.save(d)
        r2 = BM25.load(d, devices=devs)  # kwargs reach the constructor (bm25s/__init__.py:1224-1225)
        assert r2.devices == devs
        res2 = r2.retrieve(queries, k=k)
        np.testing.assert_array_equal(res2.scores, res.scores)
        monkeypatch.setenv("BM25CU_DEVICES", ",".join(str(x) for x in devs))
        r3 = BM25.load(d)
        assert r3.devices == devs
        np.testing.assert_array_equal(r3.retrieve(queries, k=k).scores, res.scores)
        monkeypatch.delenv("BM25CU_DEVICES")
