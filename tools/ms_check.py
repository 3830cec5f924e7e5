"""K1m (retrieve_ms.cu) against the streaming kernel at full size: identical ids and scores, timing, work counters.
usage: python tools/ms_check.py [shape] [k] [n_docs]   (MS_SWEEP='MS_CTAS=8,MS_PARTCOST=4096;...' runs knob sets)"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bm25s_b200._lib as L
if os.environ.get("MS_PROF", "1") == "1":
    L.SO_PATH = os.path.join(os.path.dirname(L.SO_PATH), 'libbm25cu_prof.so')
import torch
from bm25s_b200 import synth, DeviceIndex
shape = sys.argv[1] if len(sys.argv) > 1 else 'msmarco'
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n_docs, avg_len, n_vocab, n_queries = synth.SHAPES[shape]
if len(sys.argv) > 3:
    n_docs = int(sys.argv[3])
dev = torch.device('cuda')
tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev)
data, indices, indptr, nonocc = synth.csc_from_tokens_torch(tokens, ptr, n_vocab)
del tokens, ptr
qt, qp = synth.make_queries_torch(n_queries, n_vocab, seed=1, device=dev)
torch.cuda.synchronize()
import time
t0 = time.time()
ix = DeviceIndex.upload_device(data.data_ptr(), indices.data_ptr(), indptr.data_ptr(), data.numel(), n_vocab, n_docs)
print('upload_device s', round(time.time() - t0, 3))
t0 = time.time(); ix2 = DeviceIndex.upload_device(data.data_ptr(), indices.data_ptr(), indptr.data_ptr(), data.numel(), n_vocab, n_docs); print('second upload_device s', round(time.time() - t0, 3)); ix2.close(); del ix2
def run(ms):
    ix.set_option('MS', int(ms))
    oi = torch.empty((n_queries, k), dtype=torch.int32, device=dev); os_ = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
    best = None
    for it in range(4):
        st = ix.retrieve_device(qt.data_ptr(), qp.data_ptr(), n_queries, int(qp[-1]), k, oi.data_ptr(), os_.data_ptr())
        if best is None or st['fast_kernel_ms'] < best['fast_kernel_ms']:
            best = st
    return oi, os_, best
oi0, os0, st0 = run('0')
names = ['postings streamed', 'survivors queued', 'survivors probed', 'lookups', 'lookup steps', 'candidates accepted',
         'probe rounds', 'merges', 'lists streamed', 'lists skipped', 'queries served', 'queries handed on', 'dense survivors', 'dense lookups', 'dense steps', 'lookups found']
for cfg in os.environ.get('MS_SWEEP', '').split(';'):
    if not cfg:
        continue
    kv = dict(x.split('=') for x in cfg.split(','))
    kv = {kk.replace('BM25CU_', ''): int(v) for kk, v in kv.items()}
    for kk, v in kv.items():
        ix.set_option(kk, v)
    oi1, os1, st1 = run('1')
    prof = (C.c_ulonglong * 48)()
    L.check(L.lib().bm25cu_debug_prof(ix._h, prof))
    p = list(prof)[32:48]
    ok = bool((oi1 == oi0).all()) and bool((os1.view(torch.int32) == os0.view(torch.int32)).all())
    print('SWEEP', cfg, 'ms', round(st1['fast_kernel_ms'], 3), 'K1m', round(st1['ms_kernel_ms'], 3), 'identical', ok, ' '.join(f'{n.split()[0][:5]}{n.split()[-1][:6]}={v}' for n, v in zip(names, p)))
    for key in kv:
        ix.unset_option(key)
oi1, os1, st1 = run('1')
prof = (C.c_ulonglong * 48)()
L.check(L.lib().bm25cu_debug_prof(ix._h, prof))
p = list(prof)[32:48]
print('K1m+K1 ms', st1['fast_kernel_ms'], 'K1 alone ms', st0['fast_kernel_ms'], 'n_general', st1['n_general'], st0['n_general'])
same_i = bool((oi1 == oi0).all()); same_s = bool((os1.view(torch.int32) == os0.view(torch.int32)).all())
print('identical ids', same_i, 'identical scores', same_s, 'algorithmic bytes', st1['algorithmic_bytes'], st0['algorithmic_bytes'])
if not (same_i and same_s):
    bad = ((oi1 != oi0) | (os1.view(torch.int32) != os0.view(torch.int32))).any(dim=1).nonzero().flatten()
    print('differing queries', bad.numel(), bad[:10].tolist())
    for q in bad[:3].tolist():
        print(q, qt[qp[q]:qp[q + 1]].tolist()); print(oi1[q].tolist(), os1[q].tolist()); print(oi0[q].tolist(), os0[q].tolist())
for n, v in zip(names, p):
    print(f'{n:22s} {v:>14d}')
print('postings of the batch', st1['posting_bytes'] // 8)
