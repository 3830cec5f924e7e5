import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bm25s_b200._lib as L
L.SO_PATH = os.path.join(os.path.dirname(L.SO_PATH), 'libbm25cu_prof.so')
from bm25s_b200 import DeviceIndex, synth
from oracle import oracle as O, c_oracle as CO
k = 1
n_docs, n_vocab = 120_000, 20_000
corp = synth.make_corpus(n_docs, 30, n_vocab, seed=7)
data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, "bm25+")
queries = [[401, 363, 5377]]
ptr = np.array([0, 3], np.int64); flat = np.array(queries[0], np.int32)
os.environ["BM25CU_KERNEL"] = "6"; os.environ["BM25CU_TAG"] = "4096"; os.environ["BM25CU_DEBUG_DOC"] = "107091"
dev = DeviceIndex.upload(data, indices, indptr, n_docs)
ids, sc = dev.retrieve(flat, ptr, k)
rid, rsc = CO.retrieve(data, indices, indptr, n_docs, flat, ptr, k)
print(ids, sc, rid, rsc)
prof = (C.c_ulonglong * 32)()
L.check(L.lib().bm25cu_debug_prof(dev._h, prof))
p = list(prof)
print("NE probes of doc", p[29], "E resolves of doc", p[30])
for i in range(8, 16):
    x = p[i]
    if x: print(i, "kind", x >> 60, "path/kept", (x >> 56) & 0xf, "tagbyte", hex((x >> 40) & 0xff), "t", (x >> 32) & 0xff, "low", x & 0xffffffff, np.array([x & 0xffffffff], np.uint32).view(np.float32))
print("win: nmask", (p[16] >> 32) & 0xffff, "emask", p[16] & 0xffff, "wbase", p[17] >> 32, "wend", p[17] & 0xffffffff, "goff0", p[18], "n_cur[2][0]", np.int32(p[19] >> 32), "n_cur[1][0]", np.int32(p[19] & 0xffffffff),
      "cnt0", p[20] >> 32, "cnt1", p[20] & 0xffffffff, "ne0", p[21] >> 32, "theta", np.array([p[21] & 0xffffffff], np.uint32).view(np.float32), "t_cur0", p[22])
for t in queries[0]:
    print("tok", t, "beg", indptr[t], "end", indptr[t+1], "beg>>9", indptr[t] >> 9)
