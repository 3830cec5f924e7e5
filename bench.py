"""Benchmark of the bm25s query-time hot path on B200: QPS at k=10 on the MS-MARCO-shaped synthetic corpus
(8.8 M docs, 6 980 queries) - BASELINE.json's metric - with the roofline of the fused scoring + top-k kernel and the
CPU baseline (the C port of the reference algorithm, oracle/bm25_oracle.c) timed on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape msmarco|nq|scifact] [--k 10]

One "step" = one pass of the hot path over the whole query batch. N > 1 (torchrun, one rank per GPU): documents are
sharded by contiguous doc-id range, every rank scores the whole batch on its shard, local top-k lists are exchanged
with one NCCL all_gather and merged on the device (strong scaling: the corpus and the batch are fixed).
"""
import argparse
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
    ap.add_argument("--n-docs", type=int, default=0, help="override the shape's document count (debugging)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the shards' top-k rows meet - p2p: stores into peer memory over NVLink fused with the merge "
                         "(bm25cu_exchange_*), nccl: all_gather + merge kernel")
    ap.add_argument("--local-shards", action="store_true",
                    help="N > 1: every rank generates only its own doc range (seed 1000 + rank); implied by --shape 100m")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
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
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def shape_of(args):
    from bm25s_b200 import synth

    n_docs, avg_len, n_vocab, n_queries = synth.SHAPES[args.shape]
    if args.n_docs:
        n_docs = args.n_docs
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
    as `index_build_s` (not part of the metric)."""
    import torch

    from bm25s_b200 import DeviceIndex
    from bm25s_b200.sharded import ShardedIndex, shard_bounds

    n_docs, avg_len, n_vocab, n_queries = shape_of(args)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world == 1:
        dev = DeviceIndex.build_device(tokens.data_ptr(), ptr.data_ptr(), n_docs, n_vocab, args.method, args.method,
                                       1.5, 0.75, 0.5, device=local_rank)
    else:
        lo, hi = shard_bounds(n_docs, world, rank)
        if local_shards(args, world):
            t_lo, t_hi = 0, int(ptr[-1].item())
            tok_l, ptr_l = tokens, ptr
        else:
            t_lo, t_hi = int(ptr[lo].item()), int(ptr[hi].item())
            tok_l = tokens[t_lo:t_hi].contiguous()
            ptr_l = (ptr[lo:hi + 1] - t_lo).contiguous()
        sh = ShardedIndex.build(tok_l.data_ptr(), ptr_l.data_ptr(), n_vocab, method=args.method, rank=rank, world=world,
                                device=local_rank, doc_lo=lo, on_device=True, n_docs_local=hi - lo,
                                total_len_local=t_hi - t_lo)
        dev = sh.dev
    torch.cuda.synchronize()
    return dev, time.perf_counter() - t0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(csc_host, q_tok, q_ptr, n_docs, k, shift, target_seconds, n_threads=0):
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
        ids, sc = CO.retrieve(data, indices, indptr, n_docs, sub_tok, sub_ptr, k, shift=None if shift is None else shift[:n],
                              n_threads=cores)
        return time.perf_counter() - t0, ids, sc

    t_probe, _, _ = run(probe)
    per_q = t_probe / probe
    n = int(min(nq, max(probe, target_seconds / max(per_q, 1e-9))))
    t, ids, sc = run(n)
    return {"value": n / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} of {nq} queries, one pass, {t:.2f} s, threads pull queries from a shared counter"}, (n, ids, sc)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
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
              if world > 1 else "single shard"}

    import torch.distributed as dist

    if world > 1 and args.impl == "ours":
        dist.init_process_group("nccl", device_id=device)
    (tokens, ptr), (q_tok, q_ptr) = build_corpus(args, device, rank if args.impl == "ours" else 0, world if args.impl == "ours" else 1)
    dev, build_s = build_index(args, tokens, ptr, rank if args.impl == "ours" else 0, world if args.impl == "ours" else 1,
                               local_rank)
    del tokens, ptr
    torch.cuda.empty_cache()
    nnz = dev.nnz
    q_tok_h = q_tok.cpu().numpy()
    q_ptr_h = q_ptr.cpu().numpy()
    n_q_tokens = int(q_ptr_h[-1])
    shift_h = None
    csc_host = None
    if (args.impl == "reference") or (rank == 0 and world == 1 and not args.no_cpu_baseline) or dev.has_nonocc:
        data_h, indices_h, indptr_h, no_h = dev.download()
        csc_host = (data_h, indices_h, indptr_h)
        if no_h is not None:
            shift_h = np.array([no_h[q_tok_h[q_ptr_h[i]:q_ptr_h[i + 1]]].sum() for i in range(n_queries)], np.float32)

    # ------------------------------------------------------------------ reference arm: CPU port on host cores
    if args.impl == "reference":
        dev.close()
        torch.cuda.empty_cache()
        from oracle import c_oracle as CO

        cores = CO.max_threads()
        # bounded sample per step: sized from a probe so that (warmup + steps) passes end within a few minutes
        base, _ = cpu_baseline(csc_host, q_tok_h, q_ptr_h, n_docs, k, shift_h, target_seconds=2.0)
        budget = 150.0 / max(1, args.steps + args.warmup)
        n = int(max(cores, min(n_queries, base["value"] * budget)))
        sub_ptr = q_ptr_h[: n + 1]
        sub_tok = q_tok_h[: int(sub_ptr[-1])]
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            CO.retrieve(*csc_host, n_docs, sub_tok, sub_ptr, k, shift=None if shift_h is None else shift_h[:n], n_threads=cores)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        total = sum(times)
        qps = n * len(times) / total
        line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000 * total / len(times), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"first {n} of {n_queries} queries per step"},
                "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    from bm25s_b200 import _lib

    if not (rank == 0 and world == 1 and not args.no_cpu_baseline):
        csc_host = None
    stream = torch.cuda.current_stream()
    dev.set_stream(stream.cuda_stream)

    shift_d = None if shift_h is None else torch.from_numpy(shift_h).to(device)
    out_ids = torch.empty((n_queries, k), dtype=torch.int32, device=device)
    out_sc = torch.empty((n_queries, k), dtype=torch.float32, device=device)
    exch = None
    if world > 1:
        m_ids = torch.empty((n_queries, k), dtype=torch.int32, device=device)
        m_sc = torch.empty((n_queries, k), dtype=torch.float32, device=device)
        if args.exchange == "p2p":
            def all_gather_bytes(b):
                mine = torch.tensor(list(b), dtype=torch.uint8, device=device)
                allh = torch.empty((world, len(b)), dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(allh.view(-1), mine)
                return [bytes(allh[r].cpu().tolist()) for r in range(world)]

            try:
                exch = _lib.PeerExchange(local_rank, rank, world, n_queries * k, all_gather_bytes)
                ok = 1
            except Exception as e:  # CUDA IPC unavailable between these processes: say so and take the NCCL exchange
                print(f"[bench] peer-memory exchange unavailable ({e}); using all_gather + merge", file=sys.stderr)
                exch, ok = None, 0
            okt = torch.tensor([ok], device=device)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)  # every rank takes the same path; also the barrier before the first push
            torch.cuda.synchronize()
            if int(okt.item()) == 0:
                exch = None
                args.exchange = "nccl"
        if exch is None:
            g_ids = torch.empty((world, n_queries, k), dtype=torch.int32, device=device)
            g_sc = torch.empty((world, n_queries, k), dtype=torch.float32, device=device)
    stats_acc = []

    def step_device(qt, qp, collect=True):
        if world > 1 and exch is not None:
            # scoring kernels, push into the peers' memory, merge: one chain on the stream, no host round trip in between
            dev.retrieve_device_enqueue(qt.data_ptr(), qp.data_ptr(), n_queries, n_q_tokens, k, out_ids.data_ptr(), out_sc.data_ptr(),
                                        shift_ptr=None if shift_d is None else shift_d.data_ptr())
            exch.merge(out_ids.data_ptr(), out_sc.data_ptr(), n_queries, k, m_ids.data_ptr(), m_sc.data_ptr(), stream.cuda_stream)
            if not collect:
                return None
            st = dev.collect(n_queries, n_q_tokens, k)
            st["launches"] += 2
            return st
        st = dev.retrieve_device(qt.data_ptr(), qp.data_ptr(), n_queries, n_q_tokens, k, out_ids.data_ptr(), out_sc.data_ptr(),
                                 shift_ptr=None if shift_d is None else shift_d.data_ptr())
        launches = st["launches"]
        if world > 1:
            dist.all_gather_into_tensor(g_ids.view(-1), out_ids.view(-1))
            dist.all_gather_into_tensor(g_sc.view(-1), out_sc.view(-1))
            _lib.check(_lib.lib().bm25cu_topk_merge_device(
                _lib.C.c_void_p(g_ids.data_ptr()), _lib.C.c_void_p(g_sc.data_ptr()), world, n_queries, k,
                _lib.C.c_void_p(m_ids.data_ptr()), _lib.C.c_void_p(m_sc.data_ptr()), local_rank,
                _lib.C.c_void_p(stream.cuda_stream)))
            launches += 1
        st["launches"] = launches
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # kernel-path arm: queries resident in HBM
    for _ in range(args.warmup):
        step_device(q_tok, q_ptr)
    exchange_check = None
    if exch is not None:
        # untimed self-check: the peer-memory exchange returns exactly what all_gather + bm25cu_topk_merge_device return
        exch.check(stream.cuda_stream)
        g_ids = torch.empty((world, n_queries, k), dtype=torch.int32, device=device)
        g_sc = torch.empty((world, n_queries, k), dtype=torch.float32, device=device)
        c_ids, c_sc = torch.empty_like(m_ids), torch.empty_like(m_sc)
        dist.all_gather_into_tensor(g_ids.view(-1), out_ids.view(-1))
        dist.all_gather_into_tensor(g_sc.view(-1), out_sc.view(-1))
        _lib.check(_lib.lib().bm25cu_topk_merge_device(
            _lib.C.c_void_p(g_ids.data_ptr()), _lib.C.c_void_p(g_sc.data_ptr()), world, n_queries, k,
            _lib.C.c_void_p(c_ids.data_ptr()), _lib.C.c_void_p(c_sc.data_ptr()), local_rank, _lib.C.c_void_p(stream.cuda_stream)))
        torch.cuda.synchronize()
        exchange_check = bool(torch.equal(c_ids, m_ids) and torch.equal(c_sc.view(torch.int32), m_sc.view(torch.int32)))
        del g_ids, g_sc, c_ids, c_sc
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for it in range(args.steps):
        # p2p exchange: the steps are enqueued back to back, the last one collects the stats (one synchronisation)
        st_i = step_device(q_tok, q_ptr, collect=(exch is None or it == args.steps - 1))
        if st_i is not None:
            stats_acc.append(st_i)
    ev1.record()
    barrier()
    if exch is not None:
        exch.check(stream.cuda_stream)  # raises if a peer's rows never arrived
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = n_queries / (ms_per_step / 1000.0)

    # end-to-end arm: HOST buffers in, host results out, through the C ABI call a binding makes
    q_tok_pin = torch.from_numpy(q_tok_h).pin_memory()
    q_ptr_pin = torch.from_numpy(q_ptr_h).pin_memory()
    res_pin_i = torch.empty((n_queries, k), dtype=torch.int32).pin_memory()
    res_pin_s = torch.empty((n_queries, k), dtype=torch.float32).pin_memory()

    def step_e2e():
        if world == 1:
            ids, sc = dev.retrieve(q_tok_h, q_ptr_h, k, shift=shift_h)
            return ids, sc
        qt = q_tok_pin.to(device, non_blocking=True)
        qp = q_ptr_pin.to(device, non_blocking=True)
        step_device(qt, qp)
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
    h2d = n_q_tokens * 4 + (n_queries + 1) * 8 + (n_queries * 4 if shift_h is not None else 0)
    d2h = n_queries * k * 8

    # roofline of the dominant kernel (fused scoring + top-k), CUDA events inside the library on the launch stream
    fast_ms = float(np.mean([s["fast_kernel_ms"] for s in stats_acc]))
    ms_ms = float(np.mean([s.get("ms_kernel_ms", 0.0) for s in stats_acc]))
    alg_bytes = int(stats_acc[-1]["algorithmic_bytes"])
    peak, peak_src = peaks()
    achieved = alg_bytes / (fast_ms / 1000.0) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if world == 1 and args.shape == "msmarco" and k == 10 and os.path.exists(tpath):
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
            traffic = float(json.load(open(tpath))["dram_bytes_per_launch"])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": ("retrieve_ms_kernel (+ ms_plan_kernel, ms_merge_kernel; retrieve_fast_kernel serves what it hands on)"
                           if ms_ms > 0 else "retrieve_fast_kernel"),
                "kernel_ms": fast_ms, "ms_kernel_ms": ms_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "frac_of_nominal_8TBs": achieved / 8000.0}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "bm25cu_retrieve (host numpy arrays in/out)" if world == 1 else ("pinned H2D + retrieve_device + peer-memory exchange fused with the merge + D2H" if args.exchange == "p2p" else "pinned H2D + retrieve_device + all_gather + merge + D2H")},
            "gpu_launches": int(sum(s["launches"] for s in stats_acc)) if exch is None else int(stats_acc[-1]["launches"]) * args.steps,
            "roofline": roofline,
            "n_general": int(stats_acc[-1]["n_general"]), "nnz_local": int(nnz), "index_build_s": round(build_s, 3)}
    if world > 1:
        line["exchange"] = {"kind": args.exchange, "same_as_all_gather_merge": exchange_check}

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

    if rank == 0 and csc_host is not None:
        cb, (n_s, ref_ids, ref_sc) = cpu_baseline(csc_host, q_tok_h, q_ptr_h, n_docs, k, shift_h, args.cpu_seconds)
        line["cpu_baseline"] = cb
        # parity spot-check at full size on the CPU sample: scores bit-equal, ids equal above the k-th score
        ok = bool(np.array_equal(e2e_sc[:n_s], ref_sc))
        for qi in range(n_s):
            kth = ref_sc[qi, -1]
            ok &= set(e2e_ids[qi][e2e_sc[qi] > kth].tolist()) == set(ref_ids[qi][ref_sc[qi] > kth].tolist())
        line["parity_vs_cpu_port"] = {"queries": n_s, "ok": ok}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
