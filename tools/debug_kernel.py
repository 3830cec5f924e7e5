"""Debug helper: random test corpus, the kernel named by BM25CU_KERNEL_UNDER_TEST under several settings, details of mismatching queries."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bm25s_b200 import DeviceIndex, synth
from oracle import oracle as O, c_oracle as CO

k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_docs, n_vocab = 120_000, 20_000
corp = synth.make_corpus(n_docs, 30, n_vocab, seed=7)
data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "bm25+")
qs = synth.make_queries(64, n_vocab, seed=8)
extra = [[0, 1, 2], [0, 0], [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11], [5]]
queries = qs.to_lists() + extra
ptr = np.zeros(len(queries) + 1, np.int64)
np.cumsum([len(q) for q in queries], out=ptr[1:])
flat = np.array([t for q in queries for t in q], np.int32)
dev = DeviceIndex.upload(data, indices, indptr, n_docs)
rid, rsc = CO.retrieve(data, indices, indptr, n_docs, flat, ptr, k)
colmax = np.array([data[indptr[t]:indptr[t + 1]].max() if indptr[t + 1] > indptr[t] else 0 for t in range(n_vocab)])
for env in ({}, {"BM25CU_PRUNE": "0"}, {"BM25CU_HASHMAX": "0"}, {"BM25CU_TAG": "4096"}, {"BM25CU_POOL": "1024"},
            {"BM25CU_TAG": "65536", "BM25CU_HASHMAX": "0"}, {"BM25CU_THREADS": "256"}):
    os.environ["BM25CU_KERNEL"] = os.environ.get("BM25CU_KERNEL_UNDER_TEST", "7")
    for kk in ("BM25CU_PRUNE", "BM25CU_HASHMAX", "BM25CU_TAG", "BM25CU_POOL", "BM25CU_THREADS"):
        os.environ.pop(kk, None)
    os.environ.update(env)
    ids, sc = dev.retrieve(flat, ptr, k)
    bad = [q for q in range(len(queries)) if not np.array_equal(sc[q], rsc[q])]
    print("env", env, "bad queries", bad)
    for q in bad[:3]:
        j = int(np.argmax(sc[q] != rsc[q]))
        d = int(rid[q, j])
        print("  q", q, "tokens", queries[q], "rank", j, "expected doc", d, "score", rsc[q, j], "got", ids[q, j], sc[q, j])
        for t in queries[q]:
            col = indices[indptr[t]:indptr[t + 1]]
            pos = np.searchsorted(col, d)
            hit = pos < len(col) and col[pos] == d
            print("     tok", t, "df", len(col), "max", round(float(colmax[t]), 3), "contrib", float(data[indptr[t] + pos]) if hit else None,
                  "pos", int(pos), "gidx", int(indptr[t] + pos))
