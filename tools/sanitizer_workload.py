"""A small workload that walks every kernel path of the library and compares with the C oracle: the input of
`compute-sanitizer` where it is available, and of the CHECKED build (`make checked`: device-side bounds checks of every
computed index in the list-at-a-time kernel; BM25CU_LIBRARY=checked python tools/sanitizer_workload.py) where it is not."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bm25s_b200 import DeviceGroup, DeviceIndex, _lib, synth
from oracle import oracle as O, c_oracle as CO

n_docs, n_vocab = 60000, 8000
corp = synth.make_corpus(n_docs, 30, n_vocab, seed=7)
data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "bm25+")
qs = synth.make_queries(48, n_vocab, seed=8)
shift = np.array([O.query_shift(nonocc, q) for q in qs.to_lists()], np.float32)
rng = np.random.default_rng(3)
keep = rng.random(n_docs) < 0.4
dev = DeviceIndex.upload(data, indices, indptr, n_docs, nonocc=nonocc)
grp = DeviceGroup.upload(data, indices, indptr, n_docs, nonocc=nonocc, devices=[0, 0, 0])
n_calls = 0
for k in (10, 100, 300):
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, qs.tokens, qs.ptr, k, shift=shift)
    mid, msc = CO.retrieve(data, indices, indptr, n_docs, qs.tokens, qs.ptr, k, shift=shift, mask=keep.astype(np.float64))
    # default (list-at-a-time kernel, parts), many small parts, hand-over to the streaming kernel, streaming kernel alone,
    # lookups without slot tables / bitmaps, the dense general kernel, 128-thread CTAs for the shared-memory top lists
    for env in ({}, {"MS_PARTCOST": 512, "MS_MAXPARTS": 64}, {"MS_BUDGET": 64}, {"MS": 0}, {"MS_STAB": 0}, {"MS_STAB": 0, "MS_PBM": 0},
                {"MS_BTAB": 0}, {"FORCE_GENERAL": 1}, {"MS_BIG_NT": 128, "MS_BIGMERGE": 1}):
        for h in (dev, grp):
            with h.options(**env):
                ids, sc = h.retrieve(qs.tokens, qs.ptr, k, shift=shift)
                assert np.array_equal(sc, rsc), (k, env)
                kind, words = _lib.pack_weight_mask(keep)
                h.set_mask(kind, words)
                ids, sc = h.retrieve(qs.tokens, qs.ptr, k, shift=shift)
                assert np.array_equal(sc, msc), (k, env, "bitmap mask")
                h.set_mask(_lib.MASK_NONE)
                ids, sc = h.retrieve(qs.tokens, qs.ptr, k, shift=shift, mask=keep.astype(np.float64))
                assert np.array_equal(sc, msc), (k, env, "float64 mask")
                n_calls += 3
b = DeviceIndex.build(corp.tokens, corp.ptr, n_vocab, "bm25+", "bm25+", 1.5, 0.75, 0.5)
d2 = b.download()
assert np.array_equal(d2[0], data) and np.array_equal(d2[1], indices)
dense = dev.scores_dense(qs.tokens[qs.ptr[0]:qs.ptr[1]], shift=shift[0])
assert np.array_equal(dense, CO.scores_dense(data, indices, indptr, n_docs, qs.tokens[qs.ptr[0]:qs.ptr[1]], shift=shift[0]))
print(f"sanitizer workload ok: {n_calls} retrieve calls, library {os.path.basename(_lib.SO_PATH)}")
