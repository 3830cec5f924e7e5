"""Deterministic synthetic corpora / query batches of the shapes BASELINE.json names.

Document lengths are ``max(1, Poisson(L))``; tokens are i.i.d. from a shifted Zipf law
``p(r) ~ 1 / (r + 50)``, r = 1..V, drawn by inverse transform
``token = floor((V + s)^u * s^(1 - u) - s)`` with ``s = 50`` and ``u ~ U[0, 1)``; the offset stands in for
stop-word removal (the most frequent token lands in ~10 % of documents). Queries have ``1 + Poisson(3)``
tokens from the same law (duplicates allowed: they count twice, as in the reference). Token id == column.

The numpy generator is used by the tests (small shapes); the torch generator makes the same
distribution on the GPU for the benchmark shapes (hundreds of millions of tokens).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

ZIPF_SHIFT = 50.0

SHAPES = {
    # name: (n_docs, avg_doc_len, n_vocab, n_queries)
    "scifact": (5_183, 125, 30_000, 300),
    "nq": (2_680_000, 78, 1_000_000, 3_452),
    "msmarco": (8_800_000, 56, 2_000_000, 6_980),
    "100m": (100_000_000, 56, 5_000_000, 10_000),
}


@dataclass
class FlatCorpus:
    """Ragged id lists in the flat layout the C ABI takes (cf. bm25s/numba/retrieve_utils.py:114-115)."""

    tokens: np.ndarray  # int32[total]
    ptr: np.ndarray  # int64[n + 1]

    def __len__(self):
        return len(self.ptr) - 1

    def to_lists(self):
        t, p = self.tokens.tolist(), self.ptr.tolist()
        return [t[p[i]:p[i + 1]] for i in range(len(p) - 1)]


def zipf_tokens_numpy(rng: np.random.Generator, n: int, n_vocab: int) -> np.ndarray:
    u = rng.random(n)
    s = ZIPF_SHIFT
    x = np.exp(u * np.log(n_vocab + s) + (1.0 - u) * np.log(s)) - s
    return np.clip(np.floor(x), 0, n_vocab - 1).astype(np.int32)


def make_corpus(n_docs: int, avg_len: float, n_vocab: int, seed: int = 0) -> FlatCorpus:
    rng = np.random.default_rng(seed)
    lens = np.maximum(1, rng.poisson(avg_len, size=n_docs)).astype(np.int64)
    ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    return FlatCorpus(zipf_tokens_numpy(rng, int(ptr[-1]), n_vocab), ptr)


def make_queries(n_queries: int, n_vocab: int, seed: int = 1, mean_extra: float = 3.0) -> FlatCorpus:
    rng = np.random.default_rng(seed)
    lens = (1 + rng.poisson(mean_extra, size=n_queries)).astype(np.int64)
    ptr = np.zeros(n_queries + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    return FlatCorpus(zipf_tokens_numpy(rng, int(ptr[-1]), n_vocab), ptr)


def make_corpus_torch(n_docs: int, avg_len: float, n_vocab: int, seed: int = 0, device="cuda"):
    """Same law, generated on ``device`` with torch (plumbing only). Returns (tokens i32, ptr i64) tensors."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = torch.poisson(torch.full((n_docs,), float(avg_len), device=device), generator=g).clamp_(min=1).to(torch.int64)
    ptr = torch.zeros(n_docs + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=ptr[1:])
    total = int(ptr[-1].item())
    tokens = torch.empty(total, dtype=torch.int32, device=device)
    s = ZIPF_SHIFT
    chunk = 1 << 26
    import math

    la, lb = math.log(n_vocab + s), math.log(s)
    for o in range(0, total, chunk):
        n = min(chunk, total - o)
        u = torch.rand(n, device=device, dtype=torch.float64, generator=g)
        x = torch.exp(u * la + (1.0 - u) * lb) - s
        tokens[o:o + n] = x.floor_().clamp_(0, n_vocab - 1).to(torch.int32)
    torch.cuda.synchronize(device)  # the library reads these buffers on its own streams
    return tokens, ptr


def make_queries_torch(n_queries: int, n_vocab: int, seed: int = 1, mean_extra: float = 3.0, device="cuda"):
    import math

    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = (1 + torch.poisson(torch.full((n_queries,), float(mean_extra), device=device), generator=g)).to(torch.int64)
    ptr = torch.zeros(n_queries + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=ptr[1:])
    total = int(ptr[-1].item())
    s = ZIPF_SHIFT
    u = torch.rand(total, device=device, dtype=torch.float64, generator=g)
    x = torch.exp(u * math.log(n_vocab + s) + (1.0 - u) * math.log(s)) - s
    tokens = x.floor_().clamp_(0, n_vocab - 1).to(torch.int32)
    torch.cuda.synchronize(device)
    return tokens, ptr


def csc_from_tokens_torch(tokens, ptr, n_vocab: int, method: str = "lucene", k1: float = 1.5, b: float = 0.75,
                          delta: float = 0.5):
    """Benchmark/test plumbing: CSC score matrix of a flat corpus with torch ops on the corpus' device (one 64-bit
    sort). Same formulae as bm25s/scoring.py (float64 math, one float32 rounding); `torch.log` may differ from libm
    in the last float64 ulp, so use the library's GPU builder (bm25cu_index_build) where bit-parity matters.
    Returns (data f32, indices i32, indptr i64, nonocc f32 | None) tensors."""
    import torch

    n_docs = ptr.numel() - 1
    lens = ptr[1:] - ptr[:-1]
    doc = torch.repeat_interleave(torch.arange(n_docs, device=tokens.device, dtype=torch.int64), lens)
    key = tokens.to(torch.int64) * n_docs + doc
    del doc
    key, _ = torch.sort(key)
    ukey, tf = torch.unique_consecutive(key, return_counts=True)
    del key
    tok = torch.div(ukey, n_docs, rounding_mode="floor")
    d = ukey - tok * n_docs
    del ukey
    df = torch.bincount(tok, minlength=n_vocab)
    indptr = torch.zeros(n_vocab + 1, dtype=torch.int64, device=tokens.device)
    torch.cumsum(df, 0, out=indptr[1:])
    N = float(n_docs)
    dff = df.to(torch.float64)
    if method == "robertson":
        idf = torch.log(torch.clamp((N - dff + 0.5) / (dff + 0.5), min=1.0))
    elif method == "lucene":
        idf = torch.log(1 + (N - dff + 0.5) / (dff + 0.5))
    elif method == "atire":
        idf = torch.log(N / dff)
    elif method == "bm25l":
        idf = torch.log((N + 1) / (dff + 0.5))
    elif method == "bm25+":
        idf = torch.log((N + 1) / dff)
    else:
        raise ValueError(method)
    idf = torch.where(df > 0, idf, torch.zeros_like(idf))
    l_avg = lens.to(torch.float64).mean()
    l_d = lens.to(torch.float64)[d]
    tff = tf.to(torch.float64)
    norm = 1 - b + b * l_d / l_avg
    nonocc = None
    if method in ("robertson", "lucene"):
        tfc = tff / (k1 * norm + tff)
    elif method == "atire":
        tfc = (tff * (k1 + 1)) / (tff + k1 * norm)
    elif method == "bm25l":
        c = tff / norm
        tfc = ((k1 + 1) * (c + delta)) / (k1 + c + delta)
        nonocc = (idf * (((k1 + 1) * delta) / (k1 + delta))).to(torch.float32)
    else:
        tfc = ((k1 + 1) * tff) / (k1 * norm + tff) + delta
        nonocc = (idf * delta).to(torch.float32)
    score = idf.to(torch.float32).to(torch.float64)[tok] * tfc
    if nonocc is not None:
        score = score - nonocc.to(torch.float64)[tok]
    return score.to(torch.float32), d.to(torch.int32), indptr, nonocc
