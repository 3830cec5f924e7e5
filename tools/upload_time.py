import os, sys, time
sys.path.insert(0, '/root/repo')
import torch
from bm25s_b200 import synth, DeviceIndex
n_docs, avg_len, n_vocab, n_queries = synth.SHAPES['msmarco']
dev = torch.device('cuda')
tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev)
data, indices, indptr, nonocc = synth.csc_from_tokens_torch(tokens, ptr, n_vocab)
del tokens, ptr
torch.cuda.synchronize()
for i in range(3):
    t0 = time.time()
    ix = DeviceIndex.upload_device(data.data_ptr(), indices.data_ptr(), indptr.data_ptr(), data.numel(), n_vocab, n_docs)
    print('upload_device s', round(time.time() - t0, 3)); ix.close()
os.environ['BM25CU_BTAB'] = '0'; os.environ['BM25CU_PBM'] = '0'
for i in range(2):
    t0 = time.time()
    ix = DeviceIndex.upload_device(data.data_ptr(), indices.data_ptr(), indptr.data_ptr(), data.numel(), n_vocab, n_docs)
    print('no aux: upload_device s', round(time.time() - t0, 3)); ix.close()
