"""Masked retrieval at the MS-MARCO shape: fused kernel (mask in [0, 1] applied at emission) vs the dense general kernel."""
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
g = torch.Generator(device=dev); g.manual_seed(5)
mask = (torch.rand(n_docs, device=dev, generator=g) < 0.5).to(torch.float64)
def run(nq, reps=3, **env):
    for kk, v in env.items(): ix.set_option(kk, int(v))
    ms = []
    for _ in range(reps):
        st = ix.retrieve_device(qt.data_ptr(), qp.data_ptr(), nq, int(qp[nq]), k, oi.data_ptr(), os_.data_ptr(), mask_ptr=mask.data_ptr())
        ms.append(st['kernel_ms'])
    for kk in env: ix.unset_option(kk)
    return min(ms), st['n_general'], os_[:nq].clone()
m_fast, ng, s_fast = run(n_queries)
nq_small = 200
m_gen, ng2, s_gen = run(nq_small, reps=2, MASK_FAST=0)
print(f"fused kernel with a 50% 0/1 mask: {m_fast:.2f} ms per {n_queries} queries ({n_queries / m_fast * 1e3:,.0f} QPS), n_general {ng}")
print(f"dense general kernel: {m_gen:.2f} ms per {nq_small} queries ({nq_small / m_gen * 1e3:,.0f} QPS), n_general {ng2}")
print("same scores on the common queries:", bool(torch.equal(s_fast[:nq_small], s_gen)))
