"""Host -> device index upload at the MS-MARCO shape: bm25cu_index_upload from numpy arrays / memmaps."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bm25s_b200 import synth, DeviceIndex
shape = sys.argv[1] if len(sys.argv) > 1 else 'msmarco'
n_docs, avg_len, n_vocab, n_queries = synth.SHAPES[shape]
dev = torch.device('cuda')
tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev)
data, indices, indptr, nonocc = synth.csc_from_tokens_torch(tokens, ptr, n_vocab)
del tokens, ptr
h = (data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy())
del data, indices
torch.cuda.empty_cache()
nbytes = sum(a.nbytes for a in h)
for rep in range(3):
    t0 = time.perf_counter()
    ix = DeviceIndex.upload(h[0], h[1], h[2], n_docs)
    dt = time.perf_counter() - t0
    print(f"upload from numpy arrays: {dt:.3f} s, {nbytes / dt / 1e9:.2f} GB/s ({nbytes / 1e9:.2f} GB incl. finalize kernels)")
    ix.close()
d = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
names = ('data', 'indices', 'indptr')
for n, a in zip(names, h):
    np.save(os.path.join(d, n + '.npy'), a)
for rep in range(2):
    t0 = time.perf_counter()
    mm = [np.load(os.path.join(d, n + '.npy'), mmap_mode='r') for n in names]
    ix = DeviceIndex.upload(mm[0], mm[1], mm[2], n_docs)
    dt = time.perf_counter() - t0
    print(f"upload from .npy memmaps (page cache): {dt:.3f} s, {nbytes / dt / 1e9:.2f} GB/s")
    ix.close()
for n in names:
    os.remove(os.path.join(d, n + '.npy'))
os.rmdir(d)
