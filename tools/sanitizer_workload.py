import sys, numpy as np
sys.path.insert(0, '.')
from bm25s_b200 import DeviceIndex, synth
from oracle import oracle as O, c_oracle as CO
n_docs, n_vocab = 60000, 8000
corp = synth.make_corpus(n_docs, 30, n_vocab, seed=7)
data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "lucene")
qs = synth.make_queries(48, n_vocab, seed=8)
dev = DeviceIndex.upload(data, indices, indptr, n_docs)
import os
for k in (10, 300):
    # default (list-at-a-time kernel, parts), many small parts, hand-over to the streaming kernel, streaming kernel alone
    for env in ({}, {"MS_PARTCOST": 512, "MS_MAXPARTS": 64}, {"MS_BUDGET": 64}, {"MS": 0}):
        with dev.options(**env):
            ids, sc = dev.retrieve(qs.tokens, qs.ptr, k)
        rid, rsc = CO.retrieve(data, indices, indptr, n_docs, qs.tokens, qs.ptr, k)
        assert np.array_equal(sc, rsc), (k, env)
b = DeviceIndex.build(corp.tokens, corp.ptr, n_vocab, "lucene", "lucene", 1.5, 0.75, 0.5)
d2 = b.download()
assert np.array_equal(d2[0], data)
print("sanitizer workload ok")
