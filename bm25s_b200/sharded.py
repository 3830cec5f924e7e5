"""Doc-range sharding over the GPUs of one box: one process per GPU (torch.distributed), every rank
holds the CSC rows of a contiguous document range, scores the whole query batch on its shard with the fused kernel,
local top-k lists are exchanged with ONE all_gather and merged on the device (bm25cu_topk_merge). The index build
shards the same way: shard-local counting, one all_reduce of the document frequencies (+ doc count, total length),
then shard-local scoring - no posting crosses the fabric. torch is plumbing only (process group, device buffers).

The host-side protocol (bounds, all_gather layout, merge order) is backend-agnostic and is tested with gloo on CPU
(tests/test_sharding_gloo.py) with the oracle standing in for the kernels.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_docs: int, world: int, rank: int):
    """Contiguous doc-id range [lo, hi) of `rank`."""
    return n_docs * rank // world, n_docs * (rank + 1) // world


def merge_topk_numpy(ids: np.ndarray, scores: np.ndarray, k: int):
    """Reference merge of [n_parts, Q, k] partial lists -> [Q, k] (score descending, id ascending; id -1 last).
    Same order as bm25cu_topk_merge; used by the CPU tests of the protocol."""
    n_parts, Q, kk = ids.shape
    out_i = np.full((Q, k), -1, np.int32)
    out_s = np.full((Q, k), -np.inf, np.float32)
    for q in range(Q):
        i = ids[:, q, :].reshape(-1)
        s = scores[:, q, :].reshape(-1)
        valid = i >= 0
        i, s = i[valid], s[valid]
        order = np.lexsort((i, -s.astype(np.float64)))[:k]
        out_i[q, : len(order)] = i[order]
        out_s[q, : len(order)] = s[order]
    return out_i, out_s


class ShardedIndex:
    """One rank's shard + the collective retrieve. `group` is a torch.distributed process group (or None = world)."""

    def __init__(self, dev_index, n_docs_global: int, rank: int, world: int, device: int, group=None):
        self.dev = dev_index
        self.n_docs = n_docs_global
        self.rank, self.world, self.device, self.group = rank, world, device, group

    @classmethod
    def from_csc(cls, data, indices, indptr, n_docs, nonocc=None, rank=0, world=1, device=0, group=None):
        from ._lib import DeviceIndex

        lo, hi = shard_bounds(n_docs, world, rank)
        dev = DeviceIndex.upload(data, indices, indptr, n_docs, nonocc=nonocc, device=device, doc_lo=lo, doc_hi=hi)
        return cls(dev, n_docs, rank, world, device, group)

    @classmethod
    def build(cls, tokens_local, doc_ptr_local, n_vocab, method="lucene", idf_method=None, k1=1.5, b=0.75, delta=0.5,
              rank=0, world=1, device=0, group=None, doc_lo=0, on_device=False, n_docs_local=None,
              total_len_local=None):
        """GPU build of this rank's shard (bm25cu_build_begin / _finish); the document frequencies, the document count
        and the total length are all-reduced across the group in between. With `on_device`, `tokens_local` /
        `doc_ptr_local` are raw device addresses and `n_docs_local` / `total_len_local` must be given."""
        import torch
        import torch.distributed as dist

        from ._lib import Builder

        bld = Builder(tokens_local, doc_ptr_local, n_vocab, device=device, on_device=on_device, n_docs=n_docs_local)
        nd_local = bld.n_docs
        total_len = int(total_len_local) if on_device else int(np.asarray(doc_ptr_local)[-1])
        if world > 1:
            nccl = dist.get_backend(group) == "nccl"
            dev_t = torch.device("cuda", device) if nccl else torch.device("cpu")
            tdf = torch.from_numpy(bld.df().astype(np.int64)).to(dev_t)
            meta = torch.tensor([nd_local, total_len], dtype=torch.int64, device=dev_t)
            dist.all_reduce(tdf, group=group)
            dist.all_reduce(meta, group=group)
            bld.set_df(tdf.cpu().numpy().astype(np.int32))
            n_docs_global, total_global = int(meta[0].item()), int(meta[1].item())
        else:
            n_docs_global, total_global = nd_local, total_len
        dev = bld.finish(n_docs_global, total_global, method, idf_method or method, k1, b, delta, doc_lo=doc_lo)
        bld.close()
        return cls(dev, n_docs_global, rank, world, device, group)

    def retrieve(self, q_tokens, q_ptr, k, shift=None):
        """Host arrays in, merged (ids int32[Q,k], scores f32[Q,k]) out on every rank."""
        import torch
        import torch.distributed as dist

        ids, sc = self.dev.retrieve(q_tokens, q_ptr, k, shift=shift)
        if self.world == 1:
            return ids, sc
        nccl = dist.get_backend(self.group) == "nccl"
        dev_t = torch.device("cuda", self.device) if nccl else torch.device("cpu")
        ti, ts = torch.from_numpy(ids).to(dev_t), torch.from_numpy(sc).to(dev_t)
        gi = torch.empty((self.world,) + tuple(ti.shape), dtype=ti.dtype, device=dev_t)
        gs = torch.empty((self.world,) + tuple(ts.shape), dtype=ts.dtype, device=dev_t)
        dist.all_gather_into_tensor(gi.view(-1), ti.view(-1), group=self.group)
        dist.all_gather_into_tensor(gs.view(-1), ts.view(-1), group=self.group)
        from ._lib import topk_merge

        return topk_merge(gi.cpu().numpy(), gs.cpu().numpy(), device=self.device)
