"""Experiment: how fast is K1 when every query starts from its final threshold (upper bound of what better seeding buys)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bm25s_b200 import synth, DeviceIndex
n_docs, avg_len, n_vocab, n_queries = synth.SHAPES['msmarco']
dev = torch.device('cuda')
tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev)
data, indices, indptr, nonocc = synth.csc_from_tokens_torch(tokens, ptr, n_vocab)
del tokens, ptr
qt, qp = synth.make_queries_torch(n_queries, n_vocab, seed=1, device=dev)
ix = DeviceIndex.upload_device(data.data_ptr(), indices.data_ptr(), indptr.data_ptr(), data.numel(), n_vocab, n_docs)
k = 10
oi = torch.empty((n_queries, k), dtype=torch.int32, device=dev); os_ = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
def run(mode, reps=3):
    ix.set_option('DEBUG_THETA', mode)
    ms = []
    for _ in range(reps):
        st = ix.retrieve_device(qt.data_ptr(), qp.data_ptr(), n_queries, int(qp[-1]), k, oi.data_ptr(), os_.data_ptr())
        ms.append(st['fast_kernel_ms'])
    return min(ms), os_.clone()
m0, s0 = run(0)
m1, s1 = run(1)
m2, s2 = run(2)
print('normal', m0, 'writing theta', m1, 'starting from the final theta', m2, 'same scores', bool(torch.equal(s0, s2)))
