"""ctypes binding of libbm25cu.so (include/bm25cu.h). There is no CPU fallback: if the shared library is
missing, or a call fails, this module raises."""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess
import threading

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# BM25CU_LIBRARY=checked | prof selects the instrumented builds of the same sources (`make checked`, `make prof`)
SO_PATH = os.path.join(_PKG, {"checked": "libbm25cu_checked.so", "prof": "libbm25cu_prof.so"}.get(os.environ.get("BM25CU_LIBRARY", ""), "libbm25cu.so"))
CSRC = os.path.join(_PKG, "csrc")

METHOD_CODES = {"robertson": 0, "lucene": 1, "atire": 2, "bm25l": 3, "bm25+": 4}

EXPORTED_SYMBOLS = [
    "bm25cu_last_error", "bm25cu_version", "bm25cu_device_count", "bm25cu_index_upload", "bm25cu_index_upload_device", "bm25cu_index_set_stream", "bm25cu_build_begin",
    "bm25cu_build_df_device", "bm25cu_build_df_get", "bm25cu_build_df_set", "bm25cu_build_finish",
    "bm25cu_build_free", "bm25cu_index_build", "bm25cu_index_sizes", "bm25cu_index_download", "bm25cu_retrieve",
    "bm25cu_retrieve_device", "bm25cu_retrieve_device_enqueue", "bm25cu_retrieve_collect", "bm25cu_scores_dense", "bm25cu_topk_dense", "bm25cu_topk_merge",
    "bm25cu_topk_merge_device", "bm25cu_index_free", "bm25cu_index_set_option", "bm25cu_index_unset_option",
    "bm25cu_debug_colkth", "bm25cu_debug_prof", "bm25cu_index_set_mask", "bm25cu_group_set_mask", "bm25cu_index_view",
    "bm25cu_exchange_create", "bm25cu_exchange_handle_size", "bm25cu_exchange_handle", "bm25cu_exchange_connect",
    "bm25cu_exchange_merge", "bm25cu_exchange_check", "bm25cu_exchange_join", "bm25cu_exchange_free",
    "bm25cu_group_upload", "bm25cu_group_build", "bm25cu_group_sizes", "bm25cu_group_shard", "bm25cu_group_set_option",
    "bm25cu_group_unset_option", "bm25cu_group_download", "bm25cu_group_retrieve", "bm25cu_group_scores_dense",
    "bm25cu_group_free",
]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_float), ("fast_kernel_ms", C.c_float), ("total_ms", C.c_float),
        ("posting_bytes", C.c_int64), ("algorithmic_bytes", C.c_int64),
        ("n_fast", C.c_int32), ("n_general", C.c_int32), ("launches", C.c_int32), ("ms_kernel_ms", C.c_float),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ }


_lib = None
_lock = threading.Lock()


def build_extension(verbose: bool = False) -> str:
    """Compile libbm25cu.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libbm25cu.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return SO_PATH


def lib():
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(SO_PATH):
                raise ImportError(
                    f"{SO_PATH} is missing: the cuda backend has no CPU fallback. Build it with "
                    "`make -C bm25s_b200/csrc` (or `python -c 'import __graft_entry__ as g; g.build()'`)."
                )
            L = C.CDLL(SO_PATH)
            L.bm25cu_last_error.restype = C.c_char_p
            for name in EXPORTED_SYMBOLS:
                if name != "bm25cu_last_error":
                    getattr(L, name).restype = C.c_int
            _lib = L
    return _lib


def last_error() -> str:
    msg = lib().bm25cu_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int):
    """Status -> exception mapping of the ABI: 1 -> ValueError, anything else non-zero -> RuntimeError."""
    if status == 0:
        return
    msg = last_error()
    if status == 1:
        raise ValueError(msg)
    if status == 4:
        raise MemoryError(msg)
    raise RuntimeError(f"libbm25cu status {status}: {msg}")


def ptr(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


MASK_NONE, MASK_BITS, MASK_F64 = 0, 1, 2


def pack_weight_mask(mask: np.ndarray):
    """A weight mask in the form the library keeps resident (include/bm25cu.h, bm25cu_index_set_mask): 0/1 masks of any
    dtype (bool, integers, floats) become one bit per document - multiplying a score by 1 or by 0 is exact, so the bit
    carries the whole mask - anything else the exact float64 copy. Returns (kind, array)."""
    m = np.asarray(mask)
    if m.dtype == np.bool_:
        keep = m
    elif m.dtype.kind in "iuf" and (np.count_nonzero(m == 0) + np.count_nonzero(m == 1)) == m.size:
        keep = m != 0
    else:
        return MASK_F64, np.ascontiguousarray(m, dtype=np.float64)
    packed = np.packbits(keep, bitorder="little")
    words = np.zeros((len(packed) + 3) // 4 + 1, dtype=np.uint32)
    words.view(np.uint8)[: len(packed)] = packed
    return MASK_BITS, words


def device_count() -> int:
    n = C.c_int32(0)
    check(lib().bm25cu_device_count(C.byref(n)))
    return int(n.value)


class DeviceIndex:
    """Owns one bm25cu_index handle (one doc-range shard on one device)."""

    def __init__(self, handle):
        self._h = handle
        self._lock = threading.Lock()  # a handle is not thread-safe (include/bm25cu.h)
        self._opts = {}                # knobs set through set_option (the library keeps the authoritative table)
        nnz, nv, nd, lo, hn = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(lib().bm25cu_index_sizes(self._h, C.byref(nnz), C.byref(nv), C.byref(nd), C.byref(lo), C.byref(hn)))
        self.nnz, self.n_vocab, self.n_docs, self.doc_lo, self.has_nonocc = nnz.value, nv.value, nd.value, lo.value, bool(hn.value)

    @classmethod
    def upload(cls, data, indices, indptr, n_docs, nonocc=None, device=0, doc_lo=0, doc_hi=None):
        data = np.ascontiguousarray(data, dtype=np.float32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        nonocc = None if nonocc is None else np.ascontiguousarray(nonocc, dtype=np.float32)
        doc_hi = n_docs if doc_hi is None else doc_hi
        h = C.c_void_p()
        check(lib().bm25cu_index_upload(ptr(data, C.c_float), ptr(indices, C.c_int32), ptr(indptr, C.c_int64),
                                        C.c_int64(len(data)), C.c_int32(len(indptr) - 1), C.c_int32(n_docs),
                                        ptr(nonocc, C.c_float), C.c_int32(device), C.c_int32(doc_lo), C.c_int32(doc_hi),
                                        C.byref(h)))
        return cls(h)

    @classmethod
    def upload_device(cls, data_ptr, indices_ptr, indptr_ptr, nnz, n_vocab, n_docs, nonocc_ptr=None, device=0, doc_lo=0,
                      doc_hi=None):
        """CSC arrays already on the device (raw addresses, e.g. torch.Tensor.data_ptr()); copied device-to-device."""
        doc_hi = n_docs if doc_hi is None else doc_hi
        h = C.c_void_p()
        check(lib().bm25cu_index_upload_device(C.c_void_p(data_ptr), C.c_void_p(indices_ptr), C.c_void_p(indptr_ptr),
                                               C.c_int64(nnz), C.c_int32(n_vocab), C.c_int32(n_docs),
                                               C.c_void_p(nonocc_ptr), C.c_int32(device), C.c_int32(doc_lo),
                                               C.c_int32(doc_hi), C.byref(h)))
        return cls(h)

    def set_option(self, name: str, value: int):
        """Change a tuning / debugging knob of this handle (include/bm25cu.h: bm25cu_index_set_option). The BM25CU_*
        environment is only read when a handle is created."""
        with self._lock:
            check(lib().bm25cu_index_set_option(self._h, name.encode(), C.c_int32(int(value))))
        self._opts[name] = int(value)

    def unset_option(self, name: str):
        with self._lock:
            check(lib().bm25cu_index_unset_option(self._h, name.encode()))
        self._opts.pop(name, None)

    @contextlib.contextmanager
    def options(self, **kv):
        """`with dev.options(MS=0, POOL=1024): ...` - knobs set for the block, previous settings restored after it."""
        kv = {(k[7:] if k.startswith("BM25CU_") else k): int(v) for k, v in kv.items()}
        old = {k: self._opts.get(k) for k in kv}
        try:
            for k, v in kv.items():
                self.set_option(k, v)
            yield self
        finally:
            for k, v in old.items():
                if v is None:
                    self.unset_option(k)
                else:
                    self.set_option(k, v)

    def view(self) -> "DeviceIndex":
        """A second handle on the same resident shard (bm25cu_index_view): own stream and workspaces, shared index arrays;
        calls on the view and on this handle may be in flight at the same time. The view keeps this handle alive."""
        h = C.c_void_p()
        check(lib().bm25cu_index_view(self._h, C.byref(h)))
        v = DeviceIndex(h)
        v._parent = self
        return v

    def set_mask(self, kind: int, arr=None):
        """Resident weight mask of this shard: kind MASK_BITS (uint32 words, bit d%32 of word d//32 = keep local document
        d), MASK_F64 (float64[n_docs]) or MASK_NONE; applied by every retrieve call that passes no mask of its own."""
        with self._lock:
            check(lib().bm25cu_index_set_mask(self._h, C.c_int32(kind), None if arr is None else arr.ctypes.data_as(C.c_void_p)))

    def set_stream(self, stream_handle: int):
        check(lib().bm25cu_index_set_stream(self._h, C.c_void_p(stream_handle)))

    @classmethod
    def build(cls, tokens, doc_ptr, n_vocab, method, idf_method, k1, b, delta, device=0):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        doc_ptr = np.ascontiguousarray(doc_ptr, dtype=np.int64)
        h = C.c_void_p()
        check(lib().bm25cu_index_build(ptr(tokens, C.c_int32), ptr(doc_ptr, C.c_int64), C.c_int32(len(doc_ptr) - 1),
                                       C.c_int32(n_vocab), C.c_int32(METHOD_CODES[method]),
                                       C.c_int32(METHOD_CODES[idf_method]), C.c_double(k1), C.c_double(b),
                                       C.c_double(delta), C.c_int32(device), C.c_int32(0), C.byref(h)))
        return cls(h)

    @classmethod
    def build_device(cls, tokens_ptr, doc_ptr_ptr, n_docs, n_vocab, method, idf_method, k1, b, delta, device=0):
        """Corpus already on the device: int32 tokens / int64 doc_ptr raw addresses."""
        h = C.c_void_p()
        check(lib().bm25cu_index_build(C.c_void_p(tokens_ptr), C.c_void_p(doc_ptr_ptr), C.c_int32(n_docs),
                                       C.c_int32(n_vocab), C.c_int32(METHOD_CODES[method]),
                                       C.c_int32(METHOD_CODES[idf_method]), C.c_double(k1), C.c_double(b),
                                       C.c_double(delta), C.c_int32(device), C.c_int32(1), C.byref(h)))
        return cls(h)

    def download(self):
        data = np.empty(self.nnz, np.float32)
        indices = np.empty(self.nnz, np.int32)
        indptr = np.empty(self.n_vocab + 1, np.int64)
        nonocc = np.empty(self.n_vocab, np.float32) if self.has_nonocc else None
        check(lib().bm25cu_index_download(self._h, ptr(data, C.c_float), ptr(indices, C.c_int32),
                                          ptr(indptr, C.c_int64), ptr(nonocc, C.c_float)))
        return data, indices, indptr, nonocc

    def download_nonocc(self):
        """The non-occurrence array alone (float32[n_vocab]) or None."""
        if not self.has_nonocc:
            return None
        nonocc = np.empty(self.n_vocab, np.float32)
        check(lib().bm25cu_index_download(self._h, None, None, None, ptr(nonocc, C.c_float)))
        return nonocc

    def retrieve(self, q_tokens, q_ptr, k, shift=None, mask=None, with_stats=False):
        q_tokens = np.ascontiguousarray(q_tokens, dtype=np.int32)
        q_ptr = np.ascontiguousarray(q_ptr, dtype=np.int64)
        nq = len(q_ptr) - 1
        shift = None if shift is None else np.ascontiguousarray(shift, dtype=np.float32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.float64)
        ids = np.empty((nq, k), np.int32)
        scores = np.empty((nq, k), np.float32)
        st = Stats()
        with self._lock:
            check(lib().bm25cu_retrieve(self._h, ptr(q_tokens, C.c_int32), ptr(q_ptr, C.c_int64), C.c_int32(nq),
                                        C.c_int32(k), C.c_int32(1), ptr(shift, C.c_float), ptr(mask, C.c_double),
                                        ptr(ids, C.c_int32), ptr(scores, C.c_float), C.byref(st)))
        return (ids, scores, st.as_dict()) if with_stats else (ids, scores)

    def retrieve_device(self, q_tokens_ptr, q_ptr_ptr, n_queries, n_query_tokens, k, out_ids_ptr, out_scores_ptr,
                        shift_ptr=None, mask_ptr=None):
        """All pointers are raw device addresses (e.g. torch.Tensor.data_ptr()) on this handle's device."""
        st = Stats()
        with self._lock:
            check(lib().bm25cu_retrieve_device(self._h, C.c_void_p(q_tokens_ptr), C.c_void_p(q_ptr_ptr),
                                               C.c_int32(n_queries), C.c_int64(n_query_tokens), C.c_int32(k),
                                               C.c_void_p(shift_ptr), C.c_void_p(mask_ptr), C.c_void_p(out_ids_ptr),
                                               C.c_void_p(out_scores_ptr), C.byref(st)))
        return st.as_dict()

    def retrieve_device_enqueue(self, q_tokens_ptr, q_ptr_ptr, n_queries, n_query_tokens, k, out_ids_ptr, out_scores_ptr,
                                shift_ptr=None, mask_ptr=None):
        """retrieve_device without the synchronisation: the kernels are enqueued on the handle's stream; `collect` later."""
        with self._lock:
            check(lib().bm25cu_retrieve_device_enqueue(self._h, C.c_void_p(q_tokens_ptr), C.c_void_p(q_ptr_ptr),
                                                       C.c_int32(n_queries), C.c_int64(n_query_tokens), C.c_int32(k),
                                                       C.c_void_p(shift_ptr), C.c_void_p(mask_ptr), C.c_void_p(out_ids_ptr),
                                                       C.c_void_p(out_scores_ptr)))

    def collect(self, n_queries, n_query_tokens, k):
        """Synchronise the handle's stream, raise the errors of the enqueued call, return its stats."""
        st = Stats()
        with self._lock:
            check(lib().bm25cu_retrieve_collect(self._h, C.c_int32(n_queries), C.c_int64(n_query_tokens), C.c_int32(k),
                                                C.byref(st)))
        return st.as_dict()

    def scores_dense(self, q_tokens, shift=None, mask=None):
        q_tokens = np.ascontiguousarray(q_tokens, dtype=np.int32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.float64)
        out = np.empty(self.n_docs, np.float32)
        with self._lock:
            check(lib().bm25cu_scores_dense(self._h, ptr(q_tokens, C.c_int32), C.c_int32(len(q_tokens)),
                                            C.c_int32(shift is not None), C.c_float(0.0 if shift is None else float(shift)),
                                            ptr(mask, C.c_double), ptr(out, C.c_float)))
        return out

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.bm25cu_index_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceGroup:
    """Owns one bm25cu_group handle: ONE index sharded by contiguous doc-id ranges over several devices of the box, driven
    from this process (include/bm25cu.h, bm25cu_group_*). Same surface as DeviceIndex where `BM25` uses it."""

    def __init__(self, handle, devices):
        self._h = handle
        self._lock = threading.Lock()
        self._opts = {}
        self.devices = list(devices)
        nnz, nv, nd, ns, hn = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(lib().bm25cu_group_sizes(self._h, C.byref(nnz), C.byref(nv), C.byref(nd), C.byref(ns), C.byref(hn)))
        self.nnz, self.n_vocab, self.n_docs, self.n_shards, self.has_nonocc = nnz.value, nv.value, nd.value, ns.value, bool(hn.value)
        self.doc_lo = 0

    @staticmethod
    def _devs(devices):
        return np.ascontiguousarray(list(devices), dtype=np.int32)

    @classmethod
    def upload(cls, data, indices, indptr, n_docs, nonocc=None, devices=(0,)):
        data = np.ascontiguousarray(data, dtype=np.float32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        nonocc = None if nonocc is None else np.ascontiguousarray(nonocc, dtype=np.float32)
        devs = cls._devs(devices)
        h = C.c_void_p()
        check(lib().bm25cu_group_upload(ptr(data, C.c_float), ptr(indices, C.c_int32), ptr(indptr, C.c_int64), C.c_int64(len(data)),
                                        C.c_int32(len(indptr) - 1), C.c_int32(n_docs), ptr(nonocc, C.c_float),
                                        ptr(devs, C.c_int32), C.c_int32(len(devs)), C.byref(h)))
        return cls(h, devs.tolist())

    @classmethod
    def build(cls, tokens, doc_ptr, n_vocab, method, idf_method, k1, b, delta, devices=(0,)):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        doc_ptr = np.ascontiguousarray(doc_ptr, dtype=np.int64)
        devs = cls._devs(devices)
        h = C.c_void_p()
        check(lib().bm25cu_group_build(ptr(tokens, C.c_int32), ptr(doc_ptr, C.c_int64), C.c_int32(len(doc_ptr) - 1), C.c_int32(n_vocab),
                                       C.c_int32(METHOD_CODES[method]), C.c_int32(METHOD_CODES[idf_method]), C.c_double(k1),
                                       C.c_double(b), C.c_double(delta), ptr(devs, C.c_int32), C.c_int32(len(devs)), C.byref(h)))
        return cls(h, devs.tolist())

    def set_option(self, name: str, value: int):
        with self._lock:
            check(lib().bm25cu_group_set_option(self._h, name.encode(), C.c_int32(int(value))))
        self._opts[name] = int(value)

    def unset_option(self, name: str):
        with self._lock:
            check(lib().bm25cu_group_unset_option(self._h, name.encode()))
        self._opts.pop(name, None)

    options = DeviceIndex.options

    def set_mask(self, kind: int, arr=None):
        """Resident weight mask over the WHOLE index (the shards take their ranges); see DeviceIndex.set_mask."""
        with self._lock:
            check(lib().bm25cu_group_set_mask(self._h, C.c_int32(kind), None if arr is None else arr.ctypes.data_as(C.c_void_p)))

    def download(self):
        data = np.empty(self.nnz, np.float32)
        indices = np.empty(self.nnz, np.int32)
        indptr = np.empty(self.n_vocab + 1, np.int64)
        nonocc = np.empty(self.n_vocab, np.float32) if self.has_nonocc else None
        check(lib().bm25cu_group_download(self._h, ptr(data, C.c_float), ptr(indices, C.c_int32), ptr(indptr, C.c_int64),
                                          ptr(nonocc, C.c_float)))
        return data, indices, indptr, nonocc

    def retrieve(self, q_tokens, q_ptr, k, shift=None, mask=None, with_stats=False):
        q_tokens = np.ascontiguousarray(q_tokens, dtype=np.int32)
        q_ptr = np.ascontiguousarray(q_ptr, dtype=np.int64)
        nq = len(q_ptr) - 1
        shift = None if shift is None else np.ascontiguousarray(shift, dtype=np.float32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.float64)
        ids = np.empty((nq, k), np.int32)
        scores = np.empty((nq, k), np.float32)
        st = Stats()
        with self._lock:
            check(lib().bm25cu_group_retrieve(self._h, ptr(q_tokens, C.c_int32), ptr(q_ptr, C.c_int64), C.c_int32(nq), C.c_int32(k),
                                              ptr(shift, C.c_float), ptr(mask, C.c_double), ptr(ids, C.c_int32),
                                              ptr(scores, C.c_float), C.byref(st)))
        return (ids, scores, st.as_dict()) if with_stats else (ids, scores)

    def scores_dense(self, q_tokens, shift=None, mask=None):
        q_tokens = np.ascontiguousarray(q_tokens, dtype=np.int32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.float64)
        out = np.empty(self.n_docs, np.float32)
        with self._lock:
            check(lib().bm25cu_group_scores_dense(self._h, ptr(q_tokens, C.c_int32), C.c_int32(len(q_tokens)),
                                                  C.c_int32(shift is not None), C.c_float(0.0 if shift is None else float(shift)),
                                                  ptr(mask, C.c_double), ptr(out, C.c_float)))
        return out

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.bm25cu_group_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Builder:
    """Two-step GPU index build of one doc-range shard (bm25cu_build_begin / _finish). Between the two steps the
    caller sums the document frequencies, the document count and the total length over all shards."""

    def __init__(self, tokens, doc_ptr, n_vocab, device=0, on_device=False, n_docs=None):
        self._b = C.c_void_p()
        self.n_vocab = int(n_vocab)
        self.device = int(device)
        if on_device:
            self.n_docs = int(n_docs)
            check(lib().bm25cu_build_begin(C.c_void_p(tokens), C.c_void_p(doc_ptr), C.c_int32(self.n_docs),
                                           C.c_int32(n_vocab), C.c_int32(device), C.c_int32(1), C.byref(self._b)))
        else:
            tokens = np.ascontiguousarray(tokens, dtype=np.int32)
            doc_ptr = np.ascontiguousarray(doc_ptr, dtype=np.int64)
            self.n_docs = len(doc_ptr) - 1
            check(lib().bm25cu_build_begin(ptr(tokens, C.c_int32), ptr(doc_ptr, C.c_int64), C.c_int32(self.n_docs),
                                           C.c_int32(n_vocab), C.c_int32(device), C.c_int32(0), C.byref(self._b)))

    def df(self) -> np.ndarray:
        out = np.empty(self.n_vocab, np.int32)
        check(lib().bm25cu_build_df_get(self._b, ptr(out, C.c_int32)))
        return out

    def df_device_tensor(self):
        """The builder's shard-local document frequencies as a torch int32[n_vocab] CUDA tensor that ALIASES the
        library's device array (bm25cu_build_df_device): an in-place all_reduce on it is all a sharded build needs."""
        import torch

        p = C.c_void_p()
        check(lib().bm25cu_build_df_device(self._b, C.byref(p)))

        class _Arr:  # __cuda_array_interface__ view of the raw pointer
            pass

        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.n_vocab,), "typestr": "<i4", "data": (int(p.value), False), "version": 3,
                                      "strides": None}
        self._df_keepalive = a
        return torch.as_tensor(a, device=torch.device("cuda", self.device))

    def set_df(self, df: np.ndarray):
        df = np.ascontiguousarray(df, dtype=np.int32)
        assert len(df) == self.n_vocab
        check(lib().bm25cu_build_df_set(self._b, ptr(df, C.c_int32)))

    def finish(self, n_docs_global, total_len_global, method, idf_method, k1, b, delta, doc_lo=0) -> "DeviceIndex":
        h = C.c_void_p()
        check(lib().bm25cu_build_finish(self._b, C.c_int64(n_docs_global), C.c_int64(total_len_global),
                                        C.c_int32(METHOD_CODES[method]), C.c_int32(METHOD_CODES[idf_method]),
                                        C.c_double(k1), C.c_double(b), C.c_double(delta), C.c_int32(doc_lo), C.byref(h)))
        return DeviceIndex(h)

    def close(self):
        if self._b is not None and _lib is not None:
            _lib.bm25cu_build_free(self._b)
        self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def topk_dense(scores: np.ndarray, k: int, device: int = 0):
    scores = np.ascontiguousarray(scores)
    if scores.dtype == np.float64:
        is64, ct = 1, C.c_double
    else:
        scores = scores.astype(np.float32, copy=False)
        is64, ct = 0, C.c_float
    ids = np.empty(k, np.int64)
    out = np.empty(k, scores.dtype)
    check(lib().bm25cu_topk_dense(ptr(scores, ct), C.c_int32(is64), C.c_int64(scores.shape[0]), C.c_int32(k),
                                  C.c_int32(1), ptr(ids, C.c_int64), ptr(out, ct), C.c_int32(device)))
    return out, ids


def topk_merge(ids: np.ndarray, scores: np.ndarray, device: int = 0):
    """ids/scores: [n_parts, n_queries, k] -> merged [n_queries, k]."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n_parts, nq, k = ids.shape
    oi = np.empty((nq, k), np.int32)
    os_ = np.empty((nq, k), np.float32)
    check(lib().bm25cu_topk_merge(ptr(ids, C.c_int32), ptr(scores, C.c_float), C.c_int32(n_parts), C.c_int32(nq),
                                  C.c_int32(k), ptr(oi, C.c_int32), ptr(os_, C.c_float), C.c_int32(device)))
    return oi, os_


class PeerExchange:
    """Exchange + merge of the shards' top-k rows over NVLink peer memory (include/bm25cu.h, bm25cu_exchange_*): every
    rank stores its rows into every peer's buffer and merges its own copy - two kernels on the caller's stream instead of
    an all_gather and a merge kernel. `all_gather_bytes(b: bytes) -> list[bytes]` exchanges the IPC handles (rank order)."""

    def __init__(self, device: int, rank: int, world: int, max_elems: int, all_gather_bytes, barrier=None):
        self._h = C.c_void_p()
        self.rank, self.world, self.device = rank, world, device
        check(lib().bm25cu_exchange_create(C.c_int32(device), C.c_int32(rank), C.c_int32(world), C.c_int64(max_elems),
                                           C.byref(self._h)))
        hs = lib().bm25cu_exchange_handle_size()
        mine = C.create_string_buffer(hs)
        check(lib().bm25cu_exchange_handle(self._h, mine))
        handles = all_gather_bytes(mine.raw)
        assert len(handles) == world and all(len(h) == hs for h in handles)
        blob = C.create_string_buffer(b"".join(handles), hs * world)
        check(lib().bm25cu_exchange_connect(self._h, blob))
        if barrier is not None:
            barrier()  # nobody may push before everybody has opened the buffers (the caller's job if no barrier is given)

    def merge(self, ids_ptr, scores_ptr, n_queries, k, out_ids_ptr, out_scores_ptr, stream=0, shift_ptr=None):
        """All pointers are raw device addresses; enqueues two kernels on `stream` (a raw cudaStream_t), no sync.
        `shift_ptr` (float32[n_queries] on the device or None): the BM25L / BM25+ shift, added after the merge - the
        shards' rows must have been scored WITHOUT it."""
        check(lib().bm25cu_exchange_merge(self._h, C.c_void_p(ids_ptr), C.c_void_p(scores_ptr), C.c_int32(n_queries),
                                          C.c_int32(k), C.c_void_p(shift_ptr), C.c_void_p(out_ids_ptr),
                                          C.c_void_p(out_scores_ptr), C.c_void_p(stream)))

    def check(self, stream=0):
        """Make `stream` wait for the last merge, synchronise it and raise if a peer never delivered."""
        check(lib().bm25cu_exchange_check(self._h, C.c_void_p(stream)))

    def join(self, stream=0):
        """Make `stream` wait (on the device) for the merge enqueued last - it runs on a side stream of the exchange."""
        check(lib().bm25cu_exchange_join(self._h, C.c_void_p(stream)))

    def close(self):
        if self._h is not None and _lib is not None:
            lib().bm25cu_exchange_free(self._h)
            self._h = None
