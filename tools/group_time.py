import os, sys, time
sys.path.insert(0, '/root/repo')
os.environ["BM25CU_TIME"] = "1"
import numpy as np, torch
from bm25s_b200 import DeviceGroup, DeviceIndex, synth
n_docs, avg_len, n_vocab, n_queries = synth.SHAPES['msmarco']
dev_t = torch.device('cuda')
tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev_t)
ix = DeviceIndex.build_device(tokens.data_ptr(), ptr.data_ptr(), n_docs, n_vocab, "lucene", "lucene", 1.5, 0.75, 0.5)
del tokens, ptr
data, indices, indptr, _ = ix.download(); ix.close()
qt, qp = synth.make_queries_torch(n_queries, n_vocab, seed=1, device=dev_t)
qt_h, qp_h = qt.cpu().numpy(), qp.cpu().numpy()
devs = [int(x) for x in sys.argv[1].split(',')]
t0 = time.time(); grp = DeviceGroup.upload(data, indices, indptr, n_docs, devices=devs); print('upload s', time.time() - t0, file=sys.stderr)
for _ in range(5):
    t0 = time.perf_counter(); ids, sc, st = grp.retrieve(qt_h, qp_h, 10, with_stats=True); print('call ms', (time.perf_counter() - t0) * 1e3, st['kernel_ms'], file=sys.stderr)
