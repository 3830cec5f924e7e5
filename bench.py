"""Benchmark of the bm25s query-time hot path on B200: QPS at k=10 on the MS-MARCO-shaped synthetic corpus
(8.8 M docs, 6 980 queries) - BASELINE.json's metric - with the roofline of the fused scoring + top-k kernel and the
CPU baselines timed on the box's host cores: the C port of the reference algorithm (oracle/bm25_oracle.c) and, when
`baseline/_ref` holds the reference package (copied there by __graft_entry__.build()), the UNMODIFIED reference itself
(bm25s numpy and numba backends) on the same CSC.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape msmarco|nq|scifact|100m] [--k 10]

One "step" = one pass of the hot path over the whole query batch. N > 1 (torchrun, one rank per GPU): documents are
sharded by contiguous doc-id range, every rank scores the whole batch on its shard, the local top-k rows are exchanged
(stores into the peers' memory over NVLink fused with the merge, or NCCL all_gather + merge with --exchange nccl) and
merged on the device (strong scaling: the corpus and the batch are fixed). Every line carries `result_digest`, a hash of
the merged ids and score bits of the whole batch: the tie rule is deterministic (score desc, id asc), so the digest is
the same at N = 1, 2, 4, 8, and a sample of rows is checked against the C oracle on the unsharded matrix at every N.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "QPS (k=10, MSMARCO-shape 8.8M docs) at 1/2/4/8 B200; % of HBM peak"
UNIT = "queries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="msmarco")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--method", default="lucene")
    ap.add_argument("--queries", type=int, default=0, help="override the shape's batch size (BASELINE configs[3]: 10000)")
    ap.add_argument("--n-docs", type=int, default=0, help="override the shape's document count (debugging)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baselines AND the oracle spot check")
    ap.add_argument("--no-reference-backends", action="store_true", help="cpu_baseline: the C port only")
    ap.add_argument("--parity-queries", type=int, default=300, help="rows checked against the C oracle (0: none)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the shards' top-k rows meet - p2p: stores into peer memory over NVLink fused with the merge "
                         "(bm25cu_exchange_*), nccl: all_gather + merge kernel")
    ap.add_argument("--lanes", type=int, default=3,
                    help="batches in flight per GPU in the timed loop: lane 0 is the shard's handle, further lanes are views of it "
                         "(bm25cu_index_view: own stream, workspaces and exchange, shared index), steps go to the lanes round-robin; "
                         "1 = one batch at a time (also measured, as `one_batch_at_a_time`)")
    ap.add_argument("--local-shards", action="store_true",
                    help="N > 1: every rank generates only its own doc range (seed 1000 + rank); implied by --shape 100m")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe), sampled every 50 ms from before the warm-up to
    the end of the end-to-end arm, so that the short timed region is always covered by samples under load."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        # "under load": samples whose power draw is in the upper half of the observed range
        load = [s for s, p in zip(sm, pw) if pw and p >= (min(pw) + max(pw)) / 2]
        return {"sm_mhz": float(np.median(load if load else sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(load),
                "window": "warm-up + timed region + end-to-end arm, 50 ms period"}


def shape_of(args):
    from bm25s_b200 import synth

    n_docs, avg_len, n_vocab, n_queries = synth.SHAPES[args.shape]
    if args.n_docs:
        n_docs = args.n_docs
    if args.queries:
        n_queries = args.queries
    return n_docs, avg_len, n_vocab, n_queries


def local_shards(args, world):
    return world > 1 and (args.local_shards or args.shape == "100m")


def build_corpus(args, device, rank=0, world=1):
    """Synthetic corpus + queries on the GPU (setup, untimed)."""
    from bm25s_b200 import synth
    from bm25s_b200.sharded import shard_bounds

    n_docs, avg_len, n_vocab, n_queries = shape_of(args)
    if local_shards(args, world):  # corpora too large to generate whole on every rank: this rank's doc range only
        lo, hi = shard_bounds(n_docs, world, rank)
        tokens, ptr = synth.make_corpus_torch(hi - lo, avg_len, n_vocab, seed=1000 + rank, device=device)
    else:
        tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=device)
    q_tok, q_ptr = synth.make_queries_torch(n_queries, n_vocab, seed=1, device=device)
    return (tokens, ptr), (q_tok, q_ptr)


def build_index(args, tokens, ptr, rank, world, local_rank):
    """The library's own GPU index build (bm25cu_build_begin/_finish) of this rank's doc-range shard; timed, reported
    as `index_build_s` (not part of the metric). Returns a ShardedIndex (world 1: a single whole shard)."""
    import torch

    from bm25s_b200.sharded import ShardedIndex, shard_bounds

    n_docs, avg_len, n_vocab, n_queries = shape_of(args)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lo, hi = shard_bounds(n_docs, world, rank)
    if world == 1 or local_shards(args, world):
        t_lo, t_hi = 0, int(ptr[-1].item())
        tok_l, ptr_l = tokens, ptr
    else:
        t_lo, t_hi = int(ptr[lo].item()), int(ptr[hi].item())
        tok_l = tokens[t_lo:t_hi].contiguous()
        ptr_l = (ptr[lo:hi + 1] - t_lo).contiguous()
    sh = ShardedIndex.build(tok_l.data_ptr(), ptr_l.data_ptr(), n_vocab, method=args.method, rank=rank, world=world,
                            device=local_rank, doc_lo=lo, on_device=True, n_docs_local=hi - lo, total_len_local=t_hi - t_lo)
    torch.cuda.synchronize()
    return sh, time.perf_counter() - t0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def query_shifts(nonocc, q_tok, q_ptr):
    """nonoccurrence_array[ids].sum() per query by numpy itself (bm25s/__init__.py:616-617)."""
    if nonocc is None:
        return None
    return np.array([nonocc[q_tok[q_ptr[i]:q_ptr[i + 1]]].sum() for i in range(len(q_ptr) - 1)], np.float32)


def sub_batch(q_tok, q_ptr, rows):
    """The queries `rows` of a flat batch as a flat batch of their own."""
    ptr = np.zeros(len(rows) + 1, np.int64)
    np.cumsum([q_ptr[q + 1] - q_ptr[q] for q in rows], out=ptr[1:])
    tok = np.concatenate([q_tok[q_ptr[q]:q_ptr[q + 1]] for q in rows]).astype(np.int32) if len(rows) else np.zeros(0, np.int32)
    return tok, ptr


def port_baseline(csc_host, q_tok, q_ptr, n_docs, k, shift, target_seconds, n_threads=0):
    """The reference algorithm (C port, oracle/bm25_oracle.c) on the host cores over a bounded query sample."""
    from oracle import c_oracle as CO

    data, indices, indptr = csc_host
    cores = CO.max_threads() if n_threads <= 0 else n_threads
    nq = len(q_ptr) - 1
    probe = min(nq, max(cores, 8))

    def run(n):
        sub_ptr = q_ptr[: n + 1]
        sub_tok = q_tok[: int(sub_ptr[-1])]
        t0 = time.perf_counter()
        CO.retrieve(data, indices, indptr, n_docs, sub_tok, sub_ptr, k, shift=None if shift is None else shift[:n], n_threads=cores)
        return time.perf_counter() - t0

    per_q = run(probe) / probe
    n = int(min(nq, max(probe, target_seconds / max(per_q, 1e-9))))
    t = run(n)
    return {"value": n / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} of {nq} queries, one pass, {t:.2f} s, threads pull queries from a shared counter"}


def oracle_parity(csc_host, q_tok, q_ptr, n_docs, k, shift, ids, scores, n_rows):
    """Strided sample of the batch scored by the C oracle on the whole (unsharded) matrix: scores bit-equal, ids equal
    as sets above the k-th score, every returned id carries its oracle dense score (a few rows)."""
    from oracle import c_oracle as CO

    data, indices, indptr = csc_host
    nq = len(q_ptr) - 1
    rows = np.unique(np.linspace(0, nq - 1, num=min(n_rows, nq)).astype(np.int64))
    s_tok, s_ptr = sub_batch(q_tok, q_ptr, rows)
    rid, rsc = CO.retrieve(data, indices, indptr, n_docs, s_tok, s_ptr, k, shift=None if shift is None else shift[rows])
    bad = {"score_rows": 0, "id_sets": 0, "dense_scores": 0, "tail": 0}
    first_bad = None
    for j, q in enumerate(rows):
        if not np.array_equal(scores[q].view(np.int32), rsc[j].view(np.int32)):
            bad["score_rows"] += 1
            first_bad = first_bad if first_bad is not None else int(q)
        kth = rsc[j, -1]
        if set(ids[q][scores[q] > kth].tolist()) != set(rid[j][rsc[j] > kth].tolist()):
            bad["id_sets"] += 1
            first_bad = first_bad if first_bad is not None else int(q)
    dense_rows = rows[:: max(1, len(rows) // 8)]
    for q in dense_rows:
        dense = CO.scores_dense(data, indices, indptr, n_docs, q_tok[q_ptr[q]:q_ptr[q + 1]], shift=None if shift is None else shift[q])
        if not np.array_equal(dense[ids[q]], scores[q]):
            bad["dense_scores"] += 1
            first_bad = first_bad if first_bad is not None else int(q)
        rest = dense.copy()
        rest[ids[q]] = -np.inf
        if not rest.max() <= scores[q, -1]:
            bad["tail"] += 1
            first_bad = first_bad if first_bad is not None else int(q)
    ok = not any(bad.values())
    out = {"queries": int(len(rows)), "dense_rows": int(len(dense_rows)), "ok": bool(ok),
           "what": "scores bit-equal, id sets equal above the k-th score, dense-vector check of returned ids and of the tail"}
    if not ok:
        q = first_bad
        out["failures"] = bad
        out["first_bad_query"] = {"q": q, "tokens": q_tok[q_ptr[q]:q_ptr[q + 1]].tolist(), "ours_ids": ids[q].tolist(), "ours_scores": scores[q].tolist(),
                                  "oracle_ids": rid[list(rows).index(q)].tolist(), "oracle_scores": rsc[list(rows).index(q)].tolist()}
    return out


def sharded_oracle_parity(dev, rank, world, q_tok, q_ptr, k, shift, ids, scores, n_rows, threads):
    """Corpora too large to rebuild unsharded (100 M docs): every rank downloads ITS shard and scores a strided sample of the
    batch with the C oracle on it (dense vector over the shard + selection), the shards' oracle rows are gathered and merged
    on rank 0 under the library's order (score desc, id asc; shift after the merge) and compared with the GPU rows: scores
    bit-equal, id sets equal above the k-th score. SURVEY.md 8c's plan for the 100 M-doc configuration."""
    import torch.distributed as dist

    from bm25s_b200.sharded import merge_topk_numpy
    from oracle import c_oracle as CO

    data, indices, indptr, _ = dev.download()
    nq = len(q_ptr) - 1
    rows = np.unique(np.linspace(0, nq - 1, num=min(n_rows, nq)).astype(np.int64))
    s_tok, s_ptr = sub_batch(q_tok, q_ptr, rows)
    rid, rsc = CO.retrieve(data, indices, indptr, dev.n_docs, s_tok, s_ptr, k, n_threads=threads)
    rid = np.where(rid >= 0, rid + dev.doc_lo, -1).astype(np.int32)
    del data, indices, indptr
    gathered = [None] * world
    dist.all_gather_object(gathered, (rid, rsc))
    if rank != 0:
        return None
    g_ids = np.stack([g[0] for g in gathered])
    g_sc = np.stack([g[1] for g in gathered])
    m_ids, m_sc = merge_topk_numpy(g_ids, g_sc, k)
    if shift is not None:
        m_sc = (m_sc + shift[rows][:, None]).astype(np.float32)
    bad = {"score_rows": 0, "id_sets": 0}
    for j, q in enumerate(rows):
        if not np.array_equal(scores[q].view(np.int32), m_sc[j].view(np.int32)):
            bad["score_rows"] += 1
        kth = m_sc[j, -1]
        if set(ids[q][scores[q] > kth].tolist()) != set(m_ids[j][m_sc[j] > kth].tolist()):
            bad["id_sets"] += 1
    return {"queries": int(len(rows)), "ok": not any(bad.values()), "failures": bad,
            "what": "per-shard C oracle (dense vector over the downloaded shard + selection) on every rank, rows merged on rank 0 "
                    "(score desc, id asc) and compared with the GPU rows: scores bit-equal, id sets equal above the k-th score"}


def reference_module():
    """The unmodified reference package from baseline/_ref (installed there by __graft_entry__.build()), or None."""
    p = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(p, "bm25s", "__init__.py")):
        return None
    if p not in sys.path:
        sys.path.insert(0, p)
    os.environ.setdefault("DISABLE_TQDM", "1")
    try:
        import bm25s  # noqa: PLC0415

        return bm25s
    except Exception as e:  # noqa: BLE001
        print(f"[bench] baseline/_ref/bm25s does not import: {e}", file=sys.stderr)
        return None


def reference_backends(bm25s, csc_host, nonocc, n_docs, n_vocab, method, q_lists, k, seconds, ours=None):
    """The reference's own retrieve() (bm25s/__init__.py:676-939) on the same CSC, injected the way load_scores does
    (:1121-1127): numpy backend on one core, numba backend on one thread and on os.cpu_count() threads. Bounded samples.
    `ours` = (ids, scores) of the same batch for a tie-aware parity check against the numpy backend."""
    import platform

    data, indices, indptr = csc_host
    out = {"cpu_count": os.cpu_count(), "cpu": platform.processor() or platform.machine(), "numpy": np.__version__}

    def make(backend):
        r = bm25s.BM25(method=method, backend=backend)
        r.scores = {"data": data, "indices": indices, "indptr": indptr, "num_docs": n_docs}
        r.nonoccurrence_array = nonocc
        r.vocab_dict = {}
        r.unique_token_ids_set = set(range(n_vocab))
        return r

    def timed(r, qs, **kw):
        t0 = time.perf_counter()
        res = r.retrieve(qs, k=k, show_progress=False, **kw)
        return time.perf_counter() - t0, res

    # numpy backend, one core (n_threads=0): probe 2 queries, then a sample sized for `seconds`
    try:
        r = make("numpy")
        t, _ = timed(r, q_lists[:2], n_threads=0)
        n = int(max(2, min(len(q_lists), 200, seconds / max(t / 2, 1e-9))))
        t, res = timed(r, q_lists[:n], n_threads=0)
        out["numpy"] = {"value": n / t, "unit": UNIT, "cores": 1, "sample": f"first {n} queries, n_threads=0, {t:.2f} s"}
        if ours is not None:
            ids, sc = ours
            rs, ri = np.asarray(res.scores), np.asarray(res.documents)
            ok = True
            for q in range(n):
                ok &= bool(np.allclose(sc[q], rs[q], rtol=1e-5, atol=0))
                kth = rs[q, -1]
                ok &= set(ids[q][sc[q] > kth].tolist()) == set(ri[q][rs[q] > kth].tolist())
            out["numpy"]["parity_with_ours"] = {"queries": n, "ok": bool(ok), "bit_equal_scores": bool(np.array_equal(sc[:n], rs))}
    except Exception as e:  # noqa: BLE001
        out["numpy"] = {"error": repr(e)[:200]}
    try:
        import numba

        out["numba_version"] = numba.__version__
        r = make("numba")
        timed(r, q_lists[:4], n_threads=1)  # JIT compilation
        best = None
        for label, nt in (("numba_1_thread", 1), ("numba_all_threads", -1)):
            cores = 1 if nt == 1 else os.cpu_count()
            probe = max(4, cores)
            t, _ = timed(r, q_lists[:probe], n_threads=nt)
            n = int(max(probe, min(len(q_lists), seconds / max(t / probe, 1e-9))))
            t, _ = timed(r, q_lists[:n], n_threads=nt)
            out[label] = {"value": n / t, "unit": UNIT, "cores": cores, "sample": f"first {n} queries, n_threads={nt}, {t:.2f} s"}
            if best is None or out[label]["value"] > out[best]["value"]:
                best = label
        out["numba_best"] = best
    except Exception as e:  # noqa: BLE001
        out["numba_error"] = repr(e)[:200]
    return out


def digest_of(ids, scores):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(ids, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(scores, dtype=np.float32).view(np.int32).tobytes())
    return h.hexdigest()[:16]


def unsharded_csc(args, device, local_rank):
    """Rank 0, untimed: the whole (unsharded) matrix built by the library and downloaded, for the oracle check at N > 1."""
    import torch

    from bm25s_b200 import DeviceIndex, synth

    n_docs, avg_len, n_vocab, _ = shape_of(args)
    tokens, ptr = synth.make_corpus_torch(n_docs, avg_len, n_vocab, seed=0, device=device)
    full = DeviceIndex.build_device(tokens.data_ptr(), ptr.data_ptr(), n_docs, n_vocab, args.method, args.method, 1.5, 0.75, 0.5,
                                    device=local_rank)
    del tokens, ptr
    data_h, indices_h, indptr_h, no_h = full.download()
    full.close()
    torch.cuda.empty_cache()
    return (data_h, indices_h, indptr_h), no_h


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        world = 1  # the reference arm is one process on the host cores
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the product path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    n_docs, avg_len, n_vocab, n_queries = shape_of(args)
    k = args.k
    config = {"workload": f"{args.shape}-shape synthetic Zipf corpus ({n_docs} docs, avg len {avg_len}, vocab {n_vocab}), "
                          f"{n_queries} queries of 1+Poisson(3) tokens, k={k}, method={args.method}",
              "l2": "inputs larger than L2 (index 3.9 GB vs 126 MB L2), no flush between steps",
              "sharding": (f"doc-range x{world}" + (", every rank generates its own range" if local_shards(args, world) else ""))
              if world > 1 else "single shard",
              "batches_in_flight": None}

    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    (tokens, ptr), (q_tok, q_ptr) = build_corpus(args, device, rank, world)
    sh, build_s = build_index(args, tokens, ptr, rank, world, local_rank)
    dev = sh.dev
    del tokens, ptr
    torch.cuda.empty_cache()
    nnz = dev.nnz
    if nnz * 8 <= 126e6:  # the small shapes (SciFact): say what is true of them
        config["l2"] = f"index {nnz * 8 / 1e6:.1f} MB fits the 126 MB L2 and is not flushed between steps (launch-bound configuration, not a bandwidth figure)"
    elif args.shape != "msmarco" or world > 1:
        config["l2"] = f"inputs larger than L2 (this rank's postings {nnz * 8 / 1e9:.2f} GB + lookup tables vs 126 MB L2), no flush between steps"
    q_tok_h = q_tok.cpu().numpy()
    q_ptr_h = q_ptr.cpu().numpy()
    n_q_tokens = int(q_ptr_h[-1])
    want_cpu = rank == 0 and not args.no_cpu_baseline
    csc_host, no_h = None, None
    if world == 1 and (args.impl == "reference" or want_cpu):
        data_h, indices_h, indptr_h, no_h = dev.download()
        csc_host = (data_h, indices_h, indptr_h)
    elif dev.has_nonocc:
        no_h = dev.download_nonocc()
    shift_h = query_shifts(no_h, q_tok_h, q_ptr_h)

    # ------------------------------------------------------------------ reference arm: CPU implementations on host cores
    if args.impl == "reference":
        dev.close()
        torch.cuda.empty_cache()
        from oracle import c_oracle as CO

        cores = CO.max_threads()
        n_pass = max(1, args.steps + args.warmup)
        ref = reference_module()
        backends = None
        q_lists = [q_tok_h[q_ptr_h[i]:q_ptr_h[i + 1]].tolist() for i in range(n_queries)]
        if ref is not None and not args.no_reference_backends:
            backends = reference_backends(ref, csc_host, no_h, n_docs, n_vocab, args.method, q_lists, k, seconds=4.0)
        # probe of the C port, then the arm's timed passes run the FASTEST CPU implementation on this box: the unmodified
        # reference's numba backend (baseline/_ref) or the C port of the reference algorithm, each on all host threads
        base = port_baseline(csc_host, q_tok_h, q_ptr_h, n_docs, k, shift_h, target_seconds=2.0)
        arm, arm_qps = "port", base["value"]
        if backends is not None:
            for name in ("numba_1_thread", "numba_all_threads", "numpy"):
                b_ = backends.get(name)
                if isinstance(b_, dict) and b_.get("value", 0) > arm_qps:
                    arm, arm_qps = name, b_["value"]
        n = int(max(cores, min(n_queries, arm_qps * (110.0 / n_pass))))
        sub_ptr = q_ptr_h[: n + 1]
        sub_tok = q_tok_h[: int(sub_ptr[-1])]
        runner = None
        if arm != "port":
            rr = ref.BM25(method=args.method, backend="numpy" if arm == "numpy" else "numba")
            rr.scores = {"data": csc_host[0], "indices": csc_host[1], "indptr": csc_host[2], "num_docs": n_docs}
            rr.nonoccurrence_array = no_h
            rr.vocab_dict, rr.unique_token_ids_set = {}, set(range(n_vocab))
            nt = {"numba_1_thread": 1, "numba_all_threads": -1, "numpy": 0}[arm]
            runner = lambda: rr.retrieve(q_lists[:n], k=k, show_progress=False, n_threads=nt)  # noqa: E731
        else:
            runner = lambda: CO.retrieve(*csc_host, n_docs, sub_tok, sub_ptr, k, shift=None if shift_h is None else shift_h[:n], n_threads=cores)  # noqa: E731
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            runner()
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        total = sum(times)
        qps = n * len(times) / total
        arm_cores = cores if arm in ("port", "numba_all_threads") else 1
        cb = {"value": qps, "unit": UNIT, "cores": arm_cores, "kind": "port" if arm == "port" else "reference",
              "sample": f"first {n} of {n_queries} queries per step" + ("" if arm == "port" else f", unmodified bm25s ({arm}) from baseline/_ref"),
              "port": {"value": base["value"], "cores": cores, "what": "oracle/bm25_oracle.c, pthreads over queries", "sample": base["sample"]}}
        if backends is not None:
            cb["reference_backends"] = backends
        value, ms_step = qps, 1000 * total / len(times)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": cb,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "value = the fastest CPU arm on this box: the C port of the reference algorithm on all host threads, or the "
                        "unmodified reference (baseline/_ref) if one of its backends is faster; the CSC comes from the library's GPU "
                        "build (bit-identical to the reference's, tests/test_gpu_build.py), the timed region is CPU-only"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    sampler = ClockSampler(local_rank)
    if not os.environ.get("BENCH_NO_SAMPLER"):  # (debugging: is the sampler's NVML polling visible in the step time?)
        sampler.start()
    stream = torch.cuda.current_stream()
    shift_d = None if shift_h is None else torch.from_numpy(shift_h).to(device)
    work = (torch.empty((n_queries, k), dtype=torch.int32, device=device), torch.empty((n_queries, k), dtype=torch.float32, device=device))
    merged = work
    if world > 1:
        merged = (torch.empty((n_queries, k), dtype=torch.int32, device=device), torch.empty((n_queries, k), dtype=torch.float32, device=device))
        if args.exchange == "p2p":
            try:
                sh.enable_peer_exchange(n_queries, k)
                ok = 1
            except Exception as e:  # CUDA IPC unavailable between these processes: say so and take the NCCL exchange
                print(f"[bench] peer-memory exchange unavailable ({e}); using all_gather + merge", file=sys.stderr)
                ok = 0
            okt = torch.tensor([ok], device=device)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)  # every rank takes the same path
            torch.cuda.synchronize()
            if int(okt.item()) == 0:
                args.exchange = "nccl"
        if args.exchange == "nccl":
            sh.enable_nccl_exchange()

    def step_device(qt, qp):
        """The library's chain (ShardedIndex.retrieve_device_enqueue): scoring kernels, exchange, merge; no synchronisation."""
        return sh.retrieve_device_enqueue(qt, qp, k, shift=shift_d, stream=stream, out=merged if world > 1 else None, work=work)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # kernel-path arm: queries resident in HBM
    for _ in range(args.warmup):
        step_device(q_tok, q_ptr)
        sh.collect()
    exchange_check = None
    if world > 1 and args.exchange == "p2p":
        # untimed self-check: the peer-memory exchange returns exactly what all_gather + bm25cu_topk_merge_device return
        from bm25s_b200 import _lib

        g_ids = torch.empty((world, n_queries, k), dtype=torch.int32, device=device)
        g_sc = torch.empty((world, n_queries, k), dtype=torch.float32, device=device)
        c_ids, c_sc = torch.empty_like(merged[0]), torch.empty_like(merged[1])
        dist.all_gather_into_tensor(g_ids.view(-1), work[0].view(-1))
        dist.all_gather_into_tensor(g_sc.view(-1), work[1].view(-1))
        _lib.check(_lib.lib().bm25cu_topk_merge_device(
            _lib.C.c_void_p(g_ids.data_ptr()), _lib.C.c_void_p(g_sc.data_ptr()), world, n_queries, k,
            _lib.C.c_void_p(None if shift_d is None else shift_d.data_ptr()),
            _lib.C.c_void_p(c_ids.data_ptr()), _lib.C.c_void_p(c_sc.data_ptr()), local_rank, _lib.C.c_void_p(stream.cuda_stream)))
        torch.cuda.synchronize()
        exchange_check = bool(torch.equal(c_ids, merged[0]) and torch.equal(c_sc.view(torch.int32), merged[1].view(torch.int32)))
        del g_ids, g_sc, c_ids, c_sc
    # lanes: lane 0 = the shard's handle on torch's current stream; further lanes = views of it on streams of their own.
    # The views are made AFTER the one-batch-at-a-time measurements: the library sizes a batch's parts by the number of
    # handles that serve the shard (1 + live views = batches the caller keeps in flight, retrieve_ms.cu).
    n_lanes = max(1, args.lanes) if (world == 1 or args.exchange == "p2p") else 1
    lanes, lane_streams, lane_bufs = [sh], [stream], [(work, merged)]

    def make_views():
        for _ in range(n_lanes - 1):
            ln = sh.view()
            if world > 1:
                ln.enable_peer_exchange(n_queries, k)
            lw = (torch.empty_like(work[0]), torch.empty_like(work[1]))
            lm = lw if world == 1 else (torch.empty_like(merged[0]), torch.empty_like(merged[1]))
            lanes.append(ln)
            lane_streams.append(torch.cuda.Stream(device))
            lane_bufs.append((lw, lm))
        for l in range(n_lanes):  # warm up (workspaces of the views; lane 0 with the part sizes of n_lanes batches in flight)
            for _ in range(2):
                lanes[l].retrieve_device_enqueue(q_tok, q_ptr, k, shift=shift_d, stream=lane_streams[l], out=lane_bufs[l][1] if world > 1 else None,
                                                 work=lane_bufs[l][0])
                lanes[l].collect()

    def timed_loop(active):
        """K steps, round-robin over the first `active` lanes, enqueued back to back; CUDA events on torch's current stream
        around all of them (the lanes' streams fork from the first event and join before the second); max over ranks."""
        barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for l in range(1, active):
            lane_streams[l].wait_event(ev0)
        for it in range(args.steps):
            l = it % active
            ev[it][0].record(lane_streams[l])
            lanes[l].retrieve_device_enqueue(q_tok, q_ptr, k, shift=shift_d, stream=lane_streams[l], out=lane_bufs[l][1] if world > 1 else None,
                                             work=lane_bufs[l][0])
            ev[it][1].record(lane_streams[l])
        for l in range(active):
            lanes[l].join(lane_streams[l])  # N > 1: the last merges run on the exchanges' side streams; the region ends behind them
        for l in range(1, active):
            e = torch.cuda.Event()
            e.record(lane_streams[l])
            stream.wait_event(e)
        ev1.record(stream)
        barrier()
        for l in range(active):
            lanes[l].collect()  # one synchronisation per lane; errors of its last call
        t = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), float(np.mean([ev[i][0].elapsed_time(ev[i][1]) for i in range(args.steps)]))

    ms_single, step_ms_local = timed_loop(1)  # one batch at a time: per-batch latency of the device chain

    # a few separately collected steps: the library's own CUDA events around the scoring kernels (roofline), per rank
    stats_acc = []
    for _ in range(min(5, args.steps)):
        barrier()
        step_device(q_tok, q_ptr)
        stats_acc.append(sh.collect())

    # end-to-end arm: HOST buffers in, host results out
    q_tok_pin = torch.from_numpy(q_tok_h).pin_memory()
    q_ptr_pin = torch.from_numpy(q_ptr_h).pin_memory()
    res_pin_i = torch.empty((n_queries, k), dtype=torch.int32).pin_memory()
    res_pin_s = torch.empty((n_queries, k), dtype=torch.float32).pin_memory()

    def step_e2e():
        if world == 1:  # the C ABI call a binding makes: bm25cu_retrieve with host numpy arrays
            return dev.retrieve(q_tok_h, q_ptr_h, k, shift=shift_h)
        # N > 1: pinned host -> device copies, the library's sharded chain, device -> pinned host copies
        qt = q_tok_pin.to(device, non_blocking=True)
        qp = q_ptr_pin.to(device, non_blocking=True)
        m_ids, m_sc = sh.retrieve_device(qt, qp, k, shift=shift_d, stream=stream)
        res_pin_i.copy_(m_ids, non_blocking=True)
        res_pin_s.copy_(m_sc, non_blocking=True)
        torch.cuda.synchronize()
        return res_pin_i.numpy(), res_pin_s.numpy()

    for _ in range(max(1, args.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_ids, e2e_sc = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_qps = n_queries * args.steps / float(t.item())
    e2e_serial_qps, e2e_callers = e2e_qps, 1

    # several batches in flight: the views, the timed loop over all lanes
    if n_lanes > 1:
        make_views()
    ms_total = timed_loop(n_lanes)[0] if n_lanes > 1 else ms_single
    ms_per_step = ms_total / args.steps
    value = n_queries / (ms_per_step / 1000.0)
    last = (args.steps - 1) % n_lanes
    digest_device = digest_of(lane_bufs[last][1][0].cpu().numpy(), lane_bufs[last][1][1].cpu().numpy())
    if world == 1 and n_lanes > 1:
        # the same C ABI call from n_lanes host threads, one handle (the index or a view of it) each: a server with several
        # request batches in flight. Every call still copies its queries in and its rows out; K calls in all.
        import threading

        handles = [ln.dev for ln in lanes]
        share = [len(range(l, args.steps, n_lanes)) for l in range(n_lanes)]
        last_rows, failures = [None] * n_lanes, []
        gate = threading.Barrier(n_lanes + 1)

        def caller(l):
            try:
                handles[l].retrieve(q_tok_h, q_ptr_h, k, shift=shift_h)  # warm-up: the view's staging buffers
                gate.wait()
                for _ in range(share[l]):
                    last_rows[l] = handles[l].retrieve(q_tok_h, q_ptr_h, k, shift=shift_h)
            except Exception as e:  # noqa: BLE001
                failures.append(e)
                gate.abort()

        pool = [threading.Thread(target=caller, args=(l,)) for l in range(n_lanes)]
        for th in pool:
            th.start()
        gate.wait()
        t0 = time.perf_counter()
        for th in pool:
            th.join()
        e2e_par_s = time.perf_counter() - t0
        if failures:
            raise failures[0]
        for rows in last_rows:
            if rows is not None and digest_of(np.array(rows[0]), np.array(rows[1])) != digest_of(np.array(e2e_ids), np.array(e2e_sc)):
                raise RuntimeError("end-to-end arm: a concurrent caller returned different rows")
        e2e_qps, e2e_callers = n_queries * args.steps / e2e_par_s, n_lanes
    if world == 1:
        for ln in lanes[1:]:
            ln.dev.close()  # what follows (BM25.retrieve on the same resident index) runs one call at a time again
        del lanes[1:]
    clocks = sampler.stop()
    h2d = n_q_tokens * 4 + (n_queries + 1) * 8 + (n_queries * 4 if shift_h is not None and world == 1 else 0)
    d2h = n_queries * k * 8
    e2e_ids, e2e_sc = np.array(e2e_ids), np.array(e2e_sc)
    digest = digest_of(e2e_ids, e2e_sc)

    # roofline of the dominant kernel (fused scoring + top-k), CUDA events inside the library on the launch stream
    fast_ms = float(np.mean([s["fast_kernel_ms"] for s in stats_acc]))
    ms_ms = float(np.mean([s.get("ms_kernel_ms", 0.0) for s in stats_acc]))
    alg_bytes = int(stats_acc[-1]["algorithmic_bytes"])
    peak, peak_src = peaks()
    achieved = alg_bytes / (fast_ms / 1000.0) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if world == 1 and args.shape == "msmarco" and k == 10 and os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = float(tj["dram_bytes_per_launch"])
            traffic_src = ("stored constant, not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one launch from the "
                           "committed `ncu --set full` capture of this command (" + str(tj.get("source", "profiles/")) + ")")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": ("retrieve_ms_kernel (+ ms_plan_kernel, ms_merge_kernel; retrieve_fast_kernel serves what it hands on)"
                           if ms_ms > 0 else "retrieve_fast_kernel"),
                "kernel_ms": fast_ms, "ms_kernel_ms": ms_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "measured_on": "the library's CUDA events around the scoring kernels, steps run one batch at a time (kernels of "
                               "different batches do not overlap there; with batches in flight a kernel shares the GPU and its own "
                               "duration says nothing about the bytes per second of the device)",
                "with_batches_in_flight": {"achieved": alg_bytes / (ms_per_step / 1000.0) / 1e9, "frac": alg_bytes / (ms_per_step / 1000.0) / 1e9 / peak,
                                           "what": "algorithmic bytes of a batch / ms_per_step of the timed loop (everything a step runs, not the scoring kernels alone)"}}

    config["batches_in_flight"] = (f"{n_lanes}: steps go round-robin to {n_lanes} lanes (the shard's handle + views of it, bm25cu_index_view: own stream, "
                                   f"workspaces and exchange, shared index); `value` = K batches / time for all K" if n_lanes > 1 else "1")
    launches_per_step = int(stats_acc[-1]["launches"])
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "callers": e2e_callers, "one_call_at_a_time": e2e_serial_qps,
                    "api": ("bm25cu_retrieve (host numpy arrays in/out)" + (f", K calls from {e2e_callers} host threads, a handle of the same "
                            "resident index each (bm25cu_index_view)" if e2e_callers > 1 else "")) if world == 1 else
                           "pinned H2D + ShardedIndex.retrieve_device (scoring kernels + " +
                           ("peer-memory exchange fused with the merge" if args.exchange == "p2p" else "all_gather + merge") + ") + D2H"},
            "gpu_launches": launches_per_step * args.steps,
            "one_batch_at_a_time": {"value": n_queries * args.steps / (ms_single / 1000.0), "unit": UNIT, "ms_per_step": ms_single / args.steps,
                                    "what": "the same K steps on ONE lane: a batch starts when the one before it has left the GPU"},
            "roofline": roofline, "result_digest": digest, "result_digest_device_path": digest_device,
            "n_general": int(stats_acc[-1]["n_general"]), "nnz_local": int(nnz), "index_build_s": round(build_s, 3)}
    if world > 1:
        line["exchange"] = {"kind": args.exchange, "same_as_all_gather_merge": exchange_check}
        # per-rank split of a step: the scoring kernels (library events, separately collected steps) and what a step takes
        # on the rank's main stream in the timed loop (scoring + push; the merge and the wait for the slowest shard run on
        # the exchange's side stream, next to the following step's scoring)
        mine = torch.tensor([fast_ms, step_ms_local], device=device, dtype=torch.float64)
        allr = torch.empty((world, 2), device=device, dtype=torch.float64)
        dist.all_gather_into_tensor(allr.view(-1), mine)
        allr = allr.cpu().numpy()
        line["per_rank"] = {"kernel_ms": [round(float(x), 4) for x in allr[:, 0]], "main_stream_step_ms": [round(float(x), 4) for x in allr[:, 1]],
                            "kernel_ms_min": float(allr[:, 0].min()), "kernel_ms_max": float(allr[:, 0].max()),
                            "ms_per_step_minus_slowest_kernel": round(ms_per_step - float(allr[:, 0].max()), 4)}

    if rank == 0 and world == 1:
        # the user-facing call, Python lists in -> numpy arrays out (BM25.retrieve of the mirror class,
        # sharing the resident device index); host query preparation is inside the timed region
        from bm25s_b200 import BM25
        api = BM25(method=args.method)
        ph = {"data": np.zeros(1, np.float32), "indices": np.zeros(1, np.int32),
              "indptr": np.zeros(n_vocab + 1, np.int64), "num_docs": n_docs}
        api.scores, api.vocab_dict, api.unique_token_ids_set = ph, {}, set(range(n_vocab))
        api.nonoccurrence_array = no_h if shift_h is not None else None
        api._dev = dev
        api._dev_key = api._residency_key()
        q_lists = [q_tok_h[q_ptr_h[i]:q_ptr_h[i + 1]].tolist() for i in range(n_queries)]
        api.retrieve(q_lists, k=k, return_as="tuple")
        n_api = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(n_api):
            res = api.retrieve(q_lists, k=k, return_as="tuple")
        api_s = (time.perf_counter() - t0) / n_api
        line["e2e_python"] = {"value": n_queries / api_s, "unit": UNIT, "api": "BM25.retrieve(list[list[int]], k) -> Results",
                              "same_scores_as_e2e": bool(np.array_equal(res.scores, e2e_sc))}
        api._dev = None  # the handle belongs to this script

    if local_shards(args, world) and not args.no_cpu_baseline and args.parity_queries > 0:
        # every rank takes part (its own shard, its share of the host threads)
        par = sharded_oracle_parity(dev, rank, world, q_tok_h, q_ptr_h, k, shift_h, e2e_ids, e2e_sc, min(args.parity_queries, 96),
                                    threads=max(1, (os.cpu_count() or world) // world))
        if rank == 0:
            line["parity_vs_cpu_port"] = par
    if world > 1:
        barrier()  # every rank is idle from here on; ranks > 0 wait on the host for rank 0's untimed checks
    if want_cpu:
        if csc_host is None and not local_shards(args, world):
            csc_host, _ = unsharded_csc(args, device, local_rank)  # N > 1: rank 0 rebuilds the whole matrix, untimed
        if csc_host is not None:
            if args.parity_queries > 0:
                line["parity_vs_cpu_port"] = oracle_parity(csc_host, q_tok_h, q_ptr_h, n_docs, k, shift_h, e2e_ids, e2e_sc, args.parity_queries)
            if world > 1:
                # The drop-in multi-GPU form, untimed region of the torchrun job: ONE process (this rank) drives all N GPUs
                # through bm25cu_group_* - what BM25(backend="cuda", devices=[0..N-1]) calls. Host arrays in and out.
                try:
                    from bm25s_b200 import DeviceGroup

                    t0 = time.perf_counter()
                    grp = DeviceGroup.upload(*csc_host, n_docs, nonocc=no_h, devices=list(range(world)))
                    up_s = time.perf_counter() - t0
                    for _ in range(2):
                        g_ids, g_sc = grp.retrieve(q_tok_h, q_ptr_h, k, shift=shift_h)
                    n_rep = max(1, min(args.steps, 5))
                    t0 = time.perf_counter()
                    for _ in range(n_rep):
                        g_ids, g_sc = grp.retrieve(q_tok_h, q_ptr_h, k, shift=shift_h)
                    g_s = (time.perf_counter() - t0) / n_rep
                    line["e2e_single_process"] = {"value": n_queries / g_s, "unit": UNIT, "n_gpus": world, "upload_s": round(up_s, 2),
                                                  "api": "bm25cu_group_retrieve (host numpy arrays in/out; one process, one host thread per GPU, "
                                                         "peer copies to GPU 0 + merge) = BM25(backend='cuda', devices=[0..N-1]).retrieve",
                                                  "result_digest": digest_of(g_ids, g_sc), "same_rows_as_e2e": digest_of(g_ids, g_sc) == digest}
                    grp.close()
                except Exception as e:  # noqa: BLE001
                    line["e2e_single_process"] = {"error": repr(e)[:300]}
            if world == 1:
                cb = port_baseline(csc_host, q_tok_h, q_ptr_h, n_docs, k, shift_h, args.cpu_seconds)
                ref = None if args.no_reference_backends else reference_module()
                if ref is not None:
                    cb["reference_backends"] = reference_backends(ref, csc_host, no_h, n_docs, n_vocab, args.method, q_lists, k,
                                                                  seconds=max(2.0, args.cpu_seconds / 3), ours=(e2e_ids, e2e_sc))
                line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        host_side_rendezvous(dist, rank, world, "bm25_bench_rank0_done")
        dist.barrier()
        dist.destroy_process_group()
    return 0


def host_side_rendezvous(dist, rank, world, tag):
    """Ranks > 0 wait on the host (process-group store) until rank 0 has passed the same point: unlike dist.barrier() no
    kernel spins on their GPUs meanwhile, so rank 0 can measure on ALL devices of the box (the single-process group)."""
    try:
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set(tag, "1")
        else:
            store.wait([tag])
    except Exception:  # noqa: BLE001
        pass


if __name__ == "__main__":
    sys.exit(main())
