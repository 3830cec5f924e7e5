"""Masked retrieval at the MS-MARCO shape: no mask, a 50 % 0/1 mask given with every call as float64
(the reference's calling convention: 70 MB travel per call), the same mask resident as float64, resident as a bitmap
(1 bit per document, its documents dropped before any lookup), and through BM25.retrieve(weight_mask=bool array).
Prints kernel ms and end-to-end QPS (host arrays in and out) per variant, and checks that all masked variants agree."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bm25s_b200 import BM25, synth, DeviceIndex, _lib

n_docs, avg_len, n_vocab, n_queries = synth.SHAPES['msmarco']
dev_t = torch.device('cuda')
tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev_t)
ix = DeviceIndex.build_device(tokens.data_ptr(), ptr.data_ptr(), n_docs, n_vocab, "lucene", "lucene", 1.5, 0.75, 0.5)
del tokens, ptr
qt, qp = synth.make_queries_torch(n_queries, n_vocab, seed=1, device=dev_t)
qt_h, qp_h = qt.cpu().numpy(), qp.cpu().numpy()
k = 10
rng = np.random.default_rng(5)
keep = rng.random(n_docs) < 0.5
mask64 = keep.astype(np.float64)


def run(label, reps=5, **kw):
    ix.retrieve(qt_h, qp_h, k, **kw)
    t0 = time.perf_counter()
    for _ in range(reps):
        ids, sc, st = ix.retrieve(qt_h, qp_h, k, with_stats=True, **kw)
    dt = (time.perf_counter() - t0) / reps
    print(f"{label:58s} kernel {st['kernel_ms']:7.3f} ms   end to end {n_queries / dt:12,.0f} QPS   n_general {st['n_general']}")
    return sc


plain = run("no mask")
a = run("float64 mask with every call (70 MB H2D per call)", mask=mask64)
ix.set_mask(_lib.MASK_F64, mask64)
b = run("resident float64 mask")
kind, words = _lib.pack_weight_mask(keep)
ix.set_mask(kind, words)
c = run("resident bitmap (1 bit per document)")
ix.set_mask(_lib.MASK_NONE)
print("masked variants agree:", bool(np.array_equal(a, b) and np.array_equal(a, c)), " differs from unmasked:", not np.array_equal(a, plain))

# the user-facing call: BM25.retrieve(list[list[int]], weight_mask=bool array) - packed on the host, resident on the device
api = BM25(method="lucene")
ph = {"data": np.zeros(1, np.float32), "indices": np.zeros(1, np.int32), "indptr": np.zeros(n_vocab + 1, np.int64), "num_docs": n_docs}
api.scores, api.vocab_dict, api.unique_token_ids_set = ph, {}, set(range(n_vocab))
api._dev, api._dev_key = ix, api._residency_key()
q_lists = [qt_h[qp_h[i]:qp_h[i + 1]].tolist() for i in range(n_queries)]
for label, m in (("BM25.retrieve, no mask", None), ("BM25.retrieve, bool mask (writable: packed every call)", keep),
                 ("BM25.retrieve, float64 0/1 mask (writable)", mask64)):
    api.retrieve(q_lists, k=k, weight_mask=m)
    t0 = time.perf_counter()
    for _ in range(3):
        res = api.retrieve(q_lists, k=k, weight_mask=m)
    print(f"{label:58s} {n_queries / ((time.perf_counter() - t0) / 3):12,.0f} QPS", "" if m is None else f"same scores: {bool(np.array_equal(res.scores, a))}")
ro = keep.copy()
ro.setflags(write=False)
api.retrieve(q_lists, k=k, weight_mask=ro)
t0 = time.perf_counter()
for _ in range(3):
    res = api.retrieve(q_lists, k=k, weight_mask=ro)
print(f"{'BM25.retrieve, read-only bool mask (uploaded once)':58s} {n_queries / ((time.perf_counter() - t0) / 3):12,.0f} QPS", f"same scores: {bool(np.array_equal(res.scores, a))}")
api._dev = None
