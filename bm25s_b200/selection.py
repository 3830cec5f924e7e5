"""`selection.topk` on the GPU (replaces bm25s/selection.py:48-64 through bm25cu_topk_dense)."""
import numpy as np

from . import _lib


def topk(query_scores, k, backend="auto", sorted=True, device=0):
    """k largest entries of a 1-D float32/float64 vector -> (scores[k], ids int64[k]); rows come back sorted by
    score descending (for sorted=False that is one valid arbitrary order, bm25s/selection.py:33-35)."""
    if backend not in ("auto", "cuda"):
        raise ValueError("Invalid backend. bm25s_b200 only implements 'cuda' (or 'auto').")
    query_scores = np.asarray(query_scores)
    if query_scores.ndim != 1:
        raise ValueError("topk only works on a 1-dimensional array of scores")
    if query_scores.dtype not in (np.float32, np.float64):
        query_scores = query_scores.astype(np.float64 if query_scores.dtype.itemsize > 4 else np.float32)
    return _lib.topk_dense(query_scores, int(k), device=device)
