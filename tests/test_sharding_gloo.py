"""World-size-2 (and 3) CPU test of the doc-range sharding protocol over gloo: shard bounds, global ids, the
all_gather layout and the merge order. The oracle stands in for the GPU kernels (no CUDA here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _check import lists_from_flat, load_golden
from oracle import oracle as O


def canonical_topk(dense, k, base=0):
    """Top-k under the library's total order: score descending, doc id ascending; padded with (-1, -inf)."""
    order = np.lexsort((np.arange(len(dense)), -dense.astype(np.float64)))[:k]
    ids = np.full(k, -1, np.int32)
    sc = np.full(k, -np.inf, np.float32)
    ids[: len(order)] = order + base
    sc[: len(order)] = dense[order]
    return ids, sc


def shard_csc(data, indices, indptr, lo, hi):
    sel = (indices >= lo) & (indices < hi)
    col = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))[sel]
    new_ptr = np.zeros(len(indptr), np.int64)
    np.cumsum(np.bincount(col, minlength=len(indptr) - 1), out=new_ptr[1:])
    return data[sel], indices[sel] - lo, new_ptr


def _worker(rank, world, port, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bm25s_b200.sharded import merge_topk_numpy, shard_bounds

    g = load_golden("synth_medium")
    m = "bm25l"
    data, indices, indptr, nonocc = g[f"{m}_data"], g[f"{m}_indices"], g[f"{m}_indptr"], g[f"{m}_nonocc"]
    queries = lists_from_flat(g["query_tokens"], g["query_ptr"])
    N = len(g["corpus_ptr"]) - 1
    lo, hi = shard_bounds(N, world, rank)
    ld, li, lp = shard_csc(data, indices, indptr, lo, hi)
    ids = np.zeros((len(queries), k), np.int32)
    sc = np.zeros((len(queries), k), np.float32)
    for qi, q in enumerate(queries):
        dense = O.get_scores(ld, lp, li, hi - lo, q, nonocc)
        ids[qi], sc[qi] = canonical_topk(dense, k, base=lo)
    ti, ts = torch.from_numpy(ids), torch.from_numpy(sc)
    gi = torch.empty((world,) + tuple(ti.shape), dtype=ti.dtype)
    gs = torch.empty((world,) + tuple(ts.shape), dtype=ts.dtype)
    dist.all_gather_into_tensor(gi.view(-1), ti.view(-1))
    dist.all_gather_into_tensor(gs.view(-1), ts.view(-1))
    mi, ms = merge_topk_numpy(gi.numpy(), gs.numpy(), k)
    # df all-reduce of a sharded build: shard-local doc frequencies sum to the global ones
    tdf = torch.from_numpy(np.diff(lp).astype(np.int64))
    dist.all_reduce(tdf)
    ok_df = bool(np.array_equal(tdf.numpy(), np.diff(indptr)))
    if rank == 0:
        full_i = np.zeros_like(mi)
        full_s = np.zeros_like(ms)
        for qi, q in enumerate(queries):
            dense = O.get_scores(data, indptr, indices, N, q, nonocc)
            full_i[qi], full_s[qi] = canonical_topk(dense, k)
        ret["ok"] = bool(np.array_equal(mi, full_i) and np.array_equal(ms, full_s)) and ok_df
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,k", [(2, 25), (3, 1200)])
def test_doc_range_sharding_protocol_gloo(world, k):
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), k, ret), nprocs=world, join=True)
        assert ret.get("ok") is True


def test_shard_bounds_cover_everything():
    from bm25s_b200.sharded import shard_bounds

    for n in (0, 1, 7, 8_800_000, 100_000_000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
