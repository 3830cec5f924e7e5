"""Doc-range sharding over the GPUs of one box: one process per GPU (torch.distributed), every rank
holds the CSC rows of a contiguous document range, scores the whole query batch on its shard with the fused kernel,
local top-k lists are exchanged and merged on the device - `retrieve_device`: stores into the peers' memory over NVLink
fused with the merge (bm25cu_exchange_*); `retrieve`: host arrays, all_gather + bm25cu_topk_merge. The index build
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

    def view(self) -> "ShardedIndex":
        """A second lane on the same resident shard: a view of the device handle (own stream and workspaces, shared index
        arrays) wrapped as a ShardedIndex of its own - give it its own exchange (`enable_peer_exchange`, on every rank in the
        same order) and its own stream, and batches on the lanes overlap: one batch's tail, its small kernels and its
        exchange run next to the other batch's scoring."""
        return ShardedIndex(self.dev.view(), self.n_docs, self.rank, self.world, self.device, self.group)

    @classmethod
    def from_csc(cls, data, indices, indptr, n_docs, nonocc=None, rank=0, world=1, device=0, group=None):
        from ._lib import DeviceIndex

        lo, hi = shard_bounds(n_docs, world, rank)
        dev = DeviceIndex.upload(data, indices, indptr, n_docs, nonocc=nonocc, device=device, doc_lo=lo, doc_hi=hi)
        return cls(dev, n_docs, rank, world, device, group)

    @classmethod
    def build(cls, tokens_local, doc_ptr_local, n_vocab, method="lucene", idf_method=None, k1=1.5, b=0.75, delta=0.5,
              rank=0, world=1, device=0, group=None, doc_lo=None, on_device=False, n_docs_local=None,
              total_len_local=None):
        """GPU build of this rank's shard (bm25cu_build_begin / _finish); the document frequencies, the document count
        and the total length are all-reduced across the group in between. With `on_device`, `tokens_local` /
        `doc_ptr_local` are raw device addresses and `n_docs_local` / `total_len_local` must be given. `doc_lo` (global id
        of this shard's first document) defaults to the sum of the document counts of the lower ranks."""
        import torch
        import torch.distributed as dist

        from ._lib import Builder

        bld = Builder(tokens_local, doc_ptr_local, n_vocab, device=device, on_device=on_device, n_docs=n_docs_local)
        nd_local = bld.n_docs
        total_len = int(total_len_local) if on_device else int(np.asarray(doc_ptr_local)[-1])
        if world > 1:
            nccl = dist.get_backend(group) == "nccl"
            dev_t = torch.device("cuda", device) if nccl else torch.device("cpu")
            meta = torch.zeros(2 + world, dtype=torch.int64, device=dev_t)  # doc count, total length, per-rank doc counts
            meta[0], meta[1], meta[2 + rank] = nd_local, total_len, nd_local
            dist.all_reduce(meta, group=group)
            if nccl:
                # the shard-local document frequencies are summed where they lie: NCCL reduces the builder's own int32[V]
                # device array in place (bm25cu_build_df_device), no host round trip
                tdf = bld.df_device_tensor()
                torch.cuda.current_stream(dev_t).synchronize()
                dist.all_reduce(tdf, group=group)
                torch.cuda.current_stream(dev_t).synchronize()
            else:
                tdf = torch.from_numpy(bld.df().astype(np.int64))
                dist.all_reduce(tdf, group=group)
                bld.set_df(tdf.numpy().astype(np.int32))
            meta_h = meta.cpu().tolist()
            n_docs_global, total_global = int(meta_h[0]), int(meta_h[1])
            if doc_lo is None:  # contiguous ranges in rank order: exclusive scan of the shards' document counts
                doc_lo = int(sum(meta_h[2:2 + rank]))
        else:
            n_docs_global, total_global = nd_local, total_len
            if doc_lo is None:
                doc_lo = 0
        dev = bld.finish(n_docs_global, total_global, method, idf_method or method, k1, b, delta, doc_lo=doc_lo)
        bld.close()
        return cls(dev, n_docs_global, rank, world, device, group)

    def retrieve(self, q_tokens, q_ptr, k, shift=None):
        """Host arrays in, merged (ids int32[Q,k], scores f32[Q,k]) out on every rank."""
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return self.dev.retrieve(q_tokens, q_ptr, k, shift=shift)
        # shards score WITHOUT the BM25L / BM25+ shift; it is added after the merge, so the merged order (and with it the
        # rows) does not depend on the number of shards
        ids, sc = self.dev.retrieve(q_tokens, q_ptr, k)
        nccl = dist.get_backend(self.group) == "nccl"
        dev_t = torch.device("cuda", self.device) if nccl else torch.device("cpu")
        ti, ts = torch.from_numpy(ids).to(dev_t), torch.from_numpy(sc).to(dev_t)
        gi = torch.empty((self.world,) + tuple(ti.shape), dtype=ti.dtype, device=dev_t)
        gs = torch.empty((self.world,) + tuple(ts.shape), dtype=ts.dtype, device=dev_t)
        dist.all_gather_into_tensor(gi.view(-1), ti.view(-1), group=self.group)
        dist.all_gather_into_tensor(gs.view(-1), ts.view(-1), group=self.group)
        from ._lib import topk_merge

        m_ids, m_sc = topk_merge(gi.cpu().numpy(), gs.cpu().numpy(), device=self.device)
        if shift is not None:
            valid = m_ids >= 0
            m_sc = np.where(valid, (m_sc + np.asarray(shift, np.float32)[:, None]).astype(np.float32), m_sc)
        return m_ids, m_sc

    # ---- device-resident path: queries and results stay in HBM, the exchange runs over NVLink peer memory -------------
    def enable_peer_exchange(self, max_queries: int, k: int):
        """Set up the peer-memory exchange (bm25cu_exchange_*, csrc/exchange.cu) for batches of up to `max_queries`
        queries and `k` results. NCCL process group required (the IPC handles travel through one all_gather)."""
        import torch
        import torch.distributed as dist

        from ._lib import PeerExchange

        dev_t = torch.device("cuda", self.device)
        if getattr(self, "_exchange", None) is not None:  # a new size: every rank lets go of the old buffers first
            torch.cuda.synchronize(dev_t)
            dist.barrier(group=self.group)
            self._exchange.close()
            self._exchange = None

        def all_gather_bytes(b):
            mine = torch.tensor(list(b), dtype=torch.uint8, device=dev_t)
            allh = torch.empty((self.world, len(b)), dtype=torch.uint8, device=dev_t)
            dist.all_gather_into_tensor(allh.view(-1), mine, group=self.group)
            return [bytes(allh[r].cpu().tolist()) for r in range(self.world)]

        self._exchange = PeerExchange(self.device, self.rank, self.world, max_queries * k, all_gather_bytes)
        dist.barrier(group=self.group)  # nobody pushes before everybody has opened the buffers
        torch.cuda.synchronize(dev_t)
        return self._exchange

    def enable_nccl_exchange(self):
        """The plain alternative to the peer-memory exchange: NCCL all_gather of the shards' rows + bm25cu_topk_merge_device
        (what north_star describes; kept for comparison and for processes that cannot share CUDA IPC handles)."""
        self._exchange = None
        self._nccl_exchange = True

    def retrieve_device_enqueue(self, q_tokens, q_ptr, k, shift=None, stream=None, out=None, work=None):
        """The chain of `retrieve_device` without its synchronisation: scoring kernels of this shard, stores into the
        peers' memory and the merge are enqueued on `stream` (default: torch's current stream) and the merged
        (ids, scores) CUDA tensors are returned at once; call `collect` (or `join`, to stay on the device) before reading them. `out` / `work`:
        optional preallocated (ids int32[Q,k], scores float32[Q,k]) pairs for the merged and the shard-local rows."""
        import torch

        nccl = getattr(self, "_nccl_exchange", False)
        if getattr(self, "_exchange", None) is None and self.world > 1 and not nccl:
            raise RuntimeError("call enable_peer_exchange(max_queries, k) (or enable_nccl_exchange()) on every rank first")
        st = torch.cuda.current_stream() if stream is None else stream
        if getattr(self, "_stream_handle", None) != st.cuda_stream:
            self.dev.set_stream(st.cuda_stream)
            self._stream_handle = st.cuda_stream
        nq = int(q_ptr.numel()) - 1
        n_tok = int(q_tokens.numel())
        if work is None:
            work = (torch.empty((nq, k), dtype=torch.int32, device=q_ptr.device),
                    torch.empty((nq, k), dtype=torch.float32, device=q_ptr.device))
        ids, sc = work
        shift_ptr = None if shift is None else shift.data_ptr()
        # world > 1: the shard scores without the shift, the merge adds it (same rows whatever the number of shards)
        self.dev.retrieve_device_enqueue(q_tokens.data_ptr(), q_ptr.data_ptr(), nq, n_tok, k, ids.data_ptr(), sc.data_ptr(),
                                         shift_ptr=shift_ptr if self.world == 1 else None)
        self._pending = (nq, n_tok, k, st.cuda_stream)
        if self.world == 1:
            return ids, sc
        if out is None:
            out = (torch.empty_like(ids), torch.empty_like(sc))
        m_ids, m_sc = out
        if self._exchange is not None:
            self._exchange.merge(ids.data_ptr(), sc.data_ptr(), nq, k, m_ids.data_ptr(), m_sc.data_ptr(), st.cuda_stream,
                                 shift_ptr=shift_ptr)
            return m_ids, m_sc
        import torch.distributed as dist

        from . import _lib

        g_ids = torch.empty((self.world, nq, k), dtype=torch.int32, device=q_ptr.device)
        g_sc = torch.empty((self.world, nq, k), dtype=torch.float32, device=q_ptr.device)
        with torch.cuda.stream(st):
            dist.all_gather_into_tensor(g_ids.view(-1), ids.view(-1), group=self.group)
            dist.all_gather_into_tensor(g_sc.view(-1), sc.view(-1), group=self.group)
        _lib.check(_lib.lib().bm25cu_topk_merge_device(
            _lib.C.c_void_p(g_ids.data_ptr()), _lib.C.c_void_p(g_sc.data_ptr()), self.world, nq, k, _lib.C.c_void_p(shift_ptr),
            _lib.C.c_void_p(m_ids.data_ptr()), _lib.C.c_void_p(m_sc.data_ptr()), self.device, _lib.C.c_void_p(st.cuda_stream)))
        return m_ids, m_sc

    def join(self, stream=None):
        """Make `stream` (default: the stream of the last enqueue) wait on the device for the merge enqueued last: with the
        peer-memory exchange the merge runs on a side stream, so that the scoring kernels of the next batch need not wait
        for the slowest shard of this one."""
        if self.world > 1 and getattr(self, "_exchange", None) is not None and getattr(self, "_pending", None) is not None:
            self._exchange.join(self._pending[3] if stream is None else stream.cuda_stream)

    def collect(self):
        """Synchronise the chain enqueued last, raise its errors (token id out of range, a peer that never delivered)
        and return the scoring kernels' stats."""
        nq, n_tok, k, stream = self._pending
        stats = self.dev.collect(nq, n_tok, k)
        if self.world > 1 and self._exchange is not None:
            self._exchange.check(stream)
            # push + merge (k <= 32) or push + slice merge + copy-out, and the one-warp waits ahead of them (exchange.cu)
            stats["launches"] += (4 if k <= 32 else 6)
        elif self.world > 1:
            stats["launches"] += 1  # merge (the all_gather is NCCL's)
        return stats

    def retrieve_device(self, q_tokens, q_ptr, k, shift=None, stream=None):
        """torch CUDA tensors in (flat int32 token ids, int64 pointers, optional float32 shifts), merged
        (ids int32[Q,k], scores float32[Q,k]) CUDA tensors out on every rank. One chain on `stream` (default: torch's
        current stream): scoring kernels of this shard, stores into the peers' memory, merge. Returns after the
        chain has been synchronised and the call's errors have been raised."""
        res = self.retrieve_device_enqueue(q_tokens, q_ptr, k, shift=shift, stream=stream)
        self.collect()
        return res
