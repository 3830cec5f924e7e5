"""Knob sweep with batches in flight: one resident shard, L lanes (the handle + views), every setting applied to all lanes with
set_option and timed at 1 lane and at L lanes (48 batches, round-robin), rows compared with the default setting's.
usage: python tools/lanes_sweep.py [n_docs] [lanes] [k]   (SWEEP_PARTCOST=a,b,c: only those part sizes)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bm25s_b200 import synth, DeviceIndex
n_docs_full, avg_len, n_vocab, n_queries = synth.SHAPES['msmarco']
n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else n_docs_full
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device('cuda')
tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=dev)
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
qt, qp = synth.make_queries_torch(n_queries, n_vocab, seed=1, device=dev)
n_tok = int(qp[-1])
base = DeviceIndex.build_device(tokens.data_ptr(), ptr.data_ptr(), n_docs, n_vocab, "lucene", "lucene", 1.5, 0.75, 0.5)
del tokens, ptr
handles, streams, outs = [], [], []
for l in range(L):
    ix = base if l == 0 else base.view()
    st = torch.cuda.Stream()
    ix.set_stream(st.cuda_stream)
    handles.append(ix); streams.append(st)
    outs.append((torch.empty((n_queries, k), dtype=torch.int32, device=dev), torch.empty((n_queries, k), dtype=torch.float32, device=dev)))
steps = 48


def timed(active):
    best = None
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        for l in range(active):
            streams[l].wait_event(e0)
        for it in range(steps):
            l = it % active
            handles[l].retrieve_device_enqueue(qt.data_ptr(), qp.data_ptr(), n_queries, n_tok, k, outs[l][0].data_ptr(), outs[l][1].data_ptr())
        for l in range(active):
            e = torch.cuda.Event(); e.record(streams[l]); main.wait_event(e)
        e1.record(main)
        torch.cuda.synchronize()
        for l in range(active):
            handles[l].collect(n_queries, n_tok, k)
        ms = e0.elapsed_time(e1) / steps
        best = ms if best is None or ms < best else best
    return best


settings = [{}]
if os.environ.get("SWEEP_PARTCOST"):
    settings += [{"MS_PARTCOST": int(x)} for x in os.environ["SWEEP_PARTCOST"].split(",")]
    os.environ["SWEEP_ONLY"] = "1"
for pc in (8192, 16384, 32768, 65536, 131072, 262144, 1048576):
    settings.append({"MS_PARTCOST": pc})
for mp in (1, 2, 4, 8):
    settings.append({"MS_MAXPARTS": mp})
settings += [{"MS_CTAS": 8}, {"MS_CTAS": 10}, {"MS_CTAS": 12}, {"MS_PARTCOST": 131072, "MS_MAXPARTS": 8}, {"MS_PARTCOST": 16384, "MS_MAXPARTS": 64}]
if os.environ.get("SWEEP_ONLY"):
    settings = settings[:1 + len(os.environ["SWEEP_PARTCOST"].split(","))]
ref = None
for opts in settings:
    for h in handles:
        for kk, v in opts.items():
            h.set_option(kk, v)
    t1, tl = timed(1), timed(L)
    rows = (outs[0][0].clone(), outs[0][1].clone())
    if ref is None:
        ref = rows
    same = bool(torch.equal(rows[0], ref[0]) and torch.equal(rows[1], ref[1]) and all(torch.equal(outs[l][1], ref[1]) for l in range(L)))
    print(f"n_docs {n_docs} k {k} {str(opts):48s} 1 lane {t1:.4f} ms   {L} lanes {tl:.4f} ms   same rows: {same}", flush=True)
    for h in handles:
        for kk in opts:
            h.unset_option(kk)
