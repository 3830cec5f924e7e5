import os, sys, ctypes as C, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bm25s_b200._lib as L
L.SO_PATH = os.path.join(os.path.dirname(L.SO_PATH), 'libbm25cu_prof.so')
import torch
from bm25s_b200 import synth, DeviceIndex
shape = sys.argv[1] if len(sys.argv) > 1 else 'msmarco'
n_docs, avg_len, n_vocab, n_queries = synth.SHAPES[shape]
if len(sys.argv) > 3:
    n_docs = int(sys.argv[3])  # emulates one doc-range shard
dev = torch.device('cuda')
tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev)
data, indices, indptr, nonocc = synth.csc_from_tokens_torch(tokens, ptr, n_vocab)
del tokens, ptr
qt, qp = synth.make_queries_torch(n_queries, n_vocab, seed=1, device=dev)
ix = DeviceIndex.upload_device(data.data_ptr(), indices.data_ptr(), indptr.data_ptr(), data.numel(), n_vocab, n_docs)
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
oi = torch.empty((n_queries, k), dtype=torch.int32, device=dev); os_ = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
for it in range(3):
    st = ix.retrieve_device(qt.data_ptr(), qp.data_ptr(), n_queries, int(qp[-1]), k, oi.data_ptr(), os_.data_ptr())
prof = (C.c_ulonglong * 48)()
L.check(L.lib().bm25cu_debug_prof(ix._h, prof))
p = list(prof)
names = {0:'loop-top',1:'wait meta',2:'seed',3:'P1',4:'bar P1',5:'P2',6:'bar P2',7:'P3',8:'bar P3',9:'P5',10:'P4',11:'free+compact',12:'run: finalize->next',13:'run: prologue',14:'run: prod/cons body',15:'run: finalize'}
tot = sum(p[0:12])
print(st)
print('consumer cycles total (sum over CTAs, thread ctid0):', tot, 'per CTA', tot/148, 'kernel ms', st['fast_kernel_ms'], 'cycles@1.965GHz', st['fast_kernel_ms']*1.965e6)
for i in range(16):
    print(f"{i:2d} {names[i]:22s} {p[i]:>14d} {100*p[i]/max(tot,1):6.2f}%")
print('producer: loop-top', p[24], 'wait data', p[25], 'analysis+publish', p[26], 'wait free', p[27], 'issue', p[28])
print('v6: resolve pass (excl. final slow)', p[20], 'hash hits', p[18], 'overflow windows', p[19], 'slow docs (warp 0)', p[23])
print('v6: NE search', p[29], 'heavy NE windows (MN>=32K)', p[19], 'MN in them', p[21], 'windows with >=480 slow docs', p[30], 'slow docs in them', p[31])
print('sorts', p[30], 'sum P', p[31], 'spilled', p[29])
print('windows', p[16], 'ME', p[17], 'MN', p[18], 'active windows', p[19], 'keptE', p[20], 'R', p[21], 'compactions', p[22])
