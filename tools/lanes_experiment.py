"""Experiment: L handles of the same shard (the handle + views of it, bm25cu_index_view) on L streams, batches enqueued
round-robin (no synchronisation between steps): what overlapping consecutive batches on one GPU buys at a given corpus size.
usage: python tools/lanes_experiment.py [n_docs] [lanes...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bm25s_b200 import synth, DeviceIndex
n_docs_full, avg_len, n_vocab, n_queries = synth.SHAPES['msmarco']
n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else n_docs_full
lanes_list = [int(x) for x in sys.argv[2:]] or [1, 2, 3, 4]
dev = torch.device('cuda')
tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev)
k = 10
qt, qp = synth.make_queries_torch(n_queries, n_vocab, seed=1, device=dev)
n_tok = int(qp[-1])
handles, streams, outs = [], [], []
base = DeviceIndex.build_device(tokens.data_ptr(), ptr.data_ptr(), n_docs, n_vocab, "lucene", "lucene", 1.5, 0.75, 0.5)
for L in range(max(lanes_list)):
    ix = base if L == 0 else base.view()
    st = torch.cuda.Stream()
    ix.set_stream(st.cuda_stream)
    handles.append(ix); streams.append(st)
    outs.append((torch.empty((n_queries, k), dtype=torch.int32, device=dev), torch.empty((n_queries, k), dtype=torch.float32, device=dev)))
del tokens, ptr
steps = 48
for L in lanes_list:
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        for l in range(L):
            streams[l].wait_event(e0)
        for it in range(steps):
            l = it % L
            handles[l].retrieve_device_enqueue(qt.data_ptr(), qp.data_ptr(), n_queries, n_tok, k, outs[l][0].data_ptr(), outs[l][1].data_ptr())
        for l in range(L):
            e = torch.cuda.Event(); e.record(streams[l]); main.wait_event(e)
        e1.record(main)
        torch.cuda.synchronize()
        for l in range(L):
            handles[l].collect(n_queries, n_tok, k)
    ms = e0.elapsed_time(e1) / steps
    same = all(torch.equal(outs[0][1], outs[l][1]) for l in range(L))
    print(f"n_docs {n_docs} lanes {L}: {ms:.4f} ms per batch, {n_queries / ms * 1e3:,.0f} QPS, same rows on every lane: {same}")
