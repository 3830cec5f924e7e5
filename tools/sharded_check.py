"""Run under torchrun on N GPUs: ShardedIndex (doc-range shards, GPU build with the all_reduce of the document
frequencies, peer-memory exchange fused with the merge) against the C oracle on the unsharded corpus.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/sharded_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bm25s_b200 import synth
from bm25s_b200.sharded import ShardedIndex, shard_bounds
from oracle import oracle as O, c_oracle as CO

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev_t = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev_t)
n_docs, n_vocab = 150_000, 20_000
corp = synth.make_corpus(n_docs, 30, n_vocab, seed=41)
qs = synth.make_queries(300, n_vocab, seed=42)
lo, hi = shard_bounds(n_docs, world, rank)
tok_l = corp.tokens[corp.ptr[lo]:corp.ptr[hi]]
ptr_l = corp.ptr[lo:hi + 1] - corp.ptr[lo]
for method in ("lucene", "bm25+"):
    # doc_lo given for one method, derived from the lower ranks' document counts for the other
    sh = ShardedIndex.build(tok_l, ptr_l, n_vocab, method=method, rank=rank, world=world, device=local,
                            doc_lo=lo if method == "lucene" else None)
    assert sh.dev.doc_lo == lo
    data, indices, indptr, nonocc = O.build_index(corp.to_lists(), n_vocab, method)
    shift = None if nonocc is None else np.array([O.query_shift(nonocc, q) for q in qs.to_lists()], np.float32)
    qt, qp = torch.from_numpy(qs.tokens).to(dev_t), torch.from_numpy(qs.ptr).to(dev_t)
    sd = None if shift is None else torch.from_numpy(shift).to(dev_t)
    for k in (10, 100):
        sh.enable_peer_exchange(len(qs), k)
        for rep in range(3):  # both buffer parities
            ids, sc = sh.retrieve_device(qt, qp, k, shift=sd)
        rid, rsc = CO.retrieve(data, indices, indptr, n_docs, qs.tokens, qs.ptr, k, shift=shift)
        assert np.array_equal(sc.cpu().numpy(), rsc), (method, k, "scores")
        ids = ids.cpu().numpy()
        for q in range(len(qs)):
            kth = rsc[q, -1]
            assert set(ids[q][rsc[q] > kth].tolist()) == set(rid[q][rsc[q] > kth].tolist()), (method, k, q)
        hid, hsc = sh.retrieve(qs.tokens, qs.ptr, k, shift=shift)  # host arrays, NCCL all_gather + merge
        assert np.array_equal(hsc, rsc) and np.array_equal(hid, ids), (method, k, "nccl path")
    sh.enable_nccl_exchange()  # device tensors, NCCL all_gather + merge kernel instead of the peer-memory exchange
    nid, nsc = sh.retrieve_device(qt, qp, 100, shift=sd)
    assert np.array_equal(nsc.cpu().numpy(), rsc) and np.array_equal(nid.cpu().numpy(), ids), (method, "nccl device path")
    if rank == 0:
        print("sharded check ok:", method, "world", world)
dist.barrier()
dist.destroy_process_group()
