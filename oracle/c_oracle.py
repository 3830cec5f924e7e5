"""ctypes binding of oracle/bm25_oracle.c (TEST INFRASTRUCTURE ONLY; never imported by bm25s_b200)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_retrieve.restype = C.c_int
        _lib.oracle_scores_dense.restype = C.c_int
        _lib.oracle_topk.restype = C.c_int
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def scores_dense(data, indices, indptr, n_docs, q, mask=None, shift=None):
    data = np.ascontiguousarray(data, np.float32)
    indices = np.ascontiguousarray(indices, np.int32)
    indptr = np.ascontiguousarray(indptr, np.int64)
    q = np.ascontiguousarray(q, np.int32)
    m = None if mask is None else np.ascontiguousarray(mask, np.float64)
    out = np.empty(n_docs, np.float32)
    rc = lib().oracle_scores_dense(_p(data, C.c_float), _p(indices, C.c_int32), _p(indptr, C.c_int64), C.c_int64(n_docs),
                                   _p(q, C.c_int32), C.c_int64(len(q)), _p(m, C.c_double), C.c_int(shift is not None),
                                   C.c_float(0.0 if shift is None else float(shift)), _p(out, C.c_float))
    assert rc == 0
    return out


def topk(scores, k, sorted=True):
    scores = np.ascontiguousarray(scores, np.float32)
    ids = np.empty(k, np.int64)
    out = np.empty(k, np.float32)
    rc = lib().oracle_topk(_p(scores, C.c_float), C.c_int64(len(scores)), C.c_int64(k), C.c_int(bool(sorted)),
                           _p(ids, C.c_int64), _p(out, C.c_float))
    assert rc == 0
    return out, ids


def retrieve(data, indices, indptr, n_docs, q_flat, q_ptr, k, shift=None, mask=None, sorted=True, n_threads=0):
    data = np.ascontiguousarray(data, np.float32)
    indices = np.ascontiguousarray(indices, np.int32)
    indptr = np.ascontiguousarray(indptr, np.int64)
    q_flat = np.ascontiguousarray(q_flat, np.int32)
    q_ptr = np.ascontiguousarray(q_ptr, np.int64)
    nq = len(q_ptr) - 1
    sh = None if shift is None else np.ascontiguousarray(shift, np.float32)
    m = None if mask is None else np.ascontiguousarray(mask, np.float64)
    ids = np.zeros((nq, k), np.int64)
    out = np.zeros((nq, k), np.float32)
    rc = lib().oracle_retrieve(_p(data, C.c_float), _p(indices, C.c_int32), _p(indptr, C.c_int64), C.c_int64(n_docs),
                               _p(q_flat, C.c_int32), _p(q_ptr, C.c_int64), C.c_int64(nq), C.c_int64(k),
                               C.c_int(bool(sorted)), _p(sh, C.c_float), _p(m, C.c_double), _p(ids, C.c_int64),
                               _p(out, C.c_float), C.c_int(n_threads))
    assert rc == 0, rc
    return ids, out
