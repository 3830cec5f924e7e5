"""`BM25` with `backend="cuda"`: the reference's public API (bm25s/__init__.py:143-1282) over libbm25cu.so.

Same constructor, `index / retrieve / get_scores / save / load / load_scores` signatures, attributes, error
types/messages and on-disk format as the reference; the score matrix lives in HBM (uploaded lazily, keyed on the
identity of `self.scores`' arrays) and scoring + top-k run in the sm_100a kernels. There is no CPU scoring path.
"""
from __future__ import annotations

import itertools
import json
import logging
import os
import contextlib
import threading
from pathlib import Path
from typing import Any, Dict, Iterable, List, NamedTuple, Tuple, Union

import numpy as np

from . import _lib
from .corpus import JsonlCorpus, find_newline_positions, save_mmindex
from .tokenization import Tokenized, convert_tokenized_to_string_list

logger = logging.getLogger("bm25s_b200")

__version__ = "0.1.0"
_METHODS = ("robertson", "lucene", "atire", "bm25l", "bm25+")


class Results(NamedTuple):
    """(documents, scores) pair returned by `retrieve` (bm25s/__init__.py:67-87)."""

    documents: np.ndarray
    scores: np.ndarray

    def __len__(self):
        return len(self.documents)

    @classmethod
    def merge(cls, results: List["Results"]) -> "Results":
        documents = np.concatenate([r.documents for r in results], axis=0)
        scores = np.concatenate([r.scores for r in results], axis=0)
        return cls(documents=documents, scores=scores)


def is_list_of_list_of_type(obj, type_=int):
    """bm25s/__init__.py:100-118"""
    if not isinstance(obj, list) or len(obj) == 0:
        return False
    first_elem = obj[0]
    if not isinstance(first_elem, list) or len(first_elem) == 0:
        return False
    return isinstance(first_elem[0], type_)


def _is_tuple_of_list_of_tokens(obj):
    """bm25s/__init__.py:121-140"""
    if not isinstance(obj, tuple) or len(obj) == 0:
        return False
    first_elem = obj[0]
    if not isinstance(first_elem, list) or len(first_elem) == 0:
        return False
    return isinstance(first_elem[0], str)


def _flatten(id_lists) -> Tuple[np.ndarray, np.ndarray]:
    """list[list[int]] -> (flat int32, pointers int64), the layout of bm25s/numba/retrieve_utils.py:114-115."""
    n = len(id_lists)
    ptr = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.fromiter((len(q) for q in id_lists), dtype=np.int64, count=n), out=ptr[1:])
    total = int(ptr[-1])
    if total == 0:
        return np.zeros(0, dtype=np.int32), ptr
    flat = np.fromiter(itertools.chain.from_iterable(id_lists), dtype=np.int64, count=total)
    return flat, ptr


class BM25:
    def __init__(self, k1=1.5, b=0.75, delta=0.5, method="lucene", idf_method=None, dtype="float32",
                 int_dtype="int32", corpus=None, backend="cuda", csc_backend="cuda", auto_compile=True, device=0,
                 devices=None):
        """Same parameters as the reference constructor (bm25s/__init__.py:144-157). `backend` must be "cuda"
        (or "auto"); `csc_backend` and `auto_compile` are accepted for signature compatibility and ignored (the
        CSC matrix is assembled on the GPU). Extra, optional: `device` selects the CUDA device; `devices=[0, 1, ...]` (or
        the environment variable BM25CU_DEVICES="0,1,...") shards the index by contiguous doc-id ranges over several GPUs
        of the box, driven from this one process - `index / retrieve / get_scores / save / load` work unchanged."""
        self.k1 = k1
        self.b = b
        self.delta = delta
        self.dtype = dtype
        self.int_dtype = int_dtype
        self.method = method
        self.idf_method = idf_method if idf_method is not None else method
        self.methods_requiring_nonoccurrence = ("bm25l", "bm25+")
        self.corpus = corpus
        self._original_version = __version__
        self.csc_backend = csc_backend
        self.device = device
        if devices is None and os.environ.get("BM25CU_DEVICES"):
            devices = [int(x) for x in os.environ["BM25CU_DEVICES"].replace(";", ",").split(",") if x.strip() != ""]
        self.devices = None if devices is None else [int(d) for d in devices]
        if self.devices is not None:
            if len(self.devices) == 0:
                raise ValueError("`devices` must be a non-empty list of CUDA device ids")
            self.device = self.devices[0]
        if backend == "auto":
            backend = "cuda"
        if backend != "cuda":
            raise ValueError(
                f"bm25s_b200 only implements backend='cuda' (got {backend!r}); there is no numpy/numba fallback."
            )
        if np.dtype(dtype) != np.float32 or np.dtype(int_dtype) != np.int32:
            raise ValueError("the cuda backend computes in float32 scores / int32 indices (the reference defaults)")
        self.backend = backend
        self.nonoccurrence_array = None
        self._dev = None
        self._dev_key = None
        self._dev_lock = threading.Lock()
        self._mask_key = None     # what the resident device mask was made from (None: no mask resident)
        self._mask_sticky = False  # set through set_weight_mask: applies to calls that pass no mask
        self._mask_keepalive = None
        self._lanes = None          # handles of the resident index that concurrent retrieve() calls share out (_lane)
        self._lanes_cv = threading.Condition()
        _lib.lib()  # fail loudly, now, if the CUDA library is missing

    # ------------------------------------------------------------------ device residency
    def _residency_key(self):
        """What the HBM copy was made from: the array OBJECTS themselves (kept alive by the key, so an id cannot be
        reused by another array) plus the sizes. In-place edits of the arrays are not seen: call `invalidate()`."""
        sc = self.scores
        return (sc["data"], sc["indices"], sc["indptr"], sc["num_docs"], len(sc["data"]), self.nonoccurrence_array,
                tuple(self.devices) if self.devices else self.device)

    @staticmethod
    def _same_key(a, b):
        return a is not None and b is not None and len(a) == len(b) and all(
            (x is y) if isinstance(x, np.ndarray) or x is None or y is None else (x == y) for x, y in zip(a, b))

    def invalidate(self):
        """Drop the HBM copy of `self.scores`; the next retrieve / get_scores uploads again. Call it after editing
        `self.scores[...]` or `nonoccurrence_array` in place (swapping the arrays is detected without it)."""
        with self._dev_lock:
            self._close_dev()
            self._dev_key = None
            if not self._mask_sticky:
                self._mask_key = None

    def _close_dev(self):
        """Free the device copy: its views first (include/bm25cu.h, bm25cu_index_view), then the handle itself."""
        with self._lanes_cv:
            st, self._lanes = self._lanes, None
        if st is not None:
            for v in st["views"]:
                v.close()
        if self._dev is not None:
            self._dev.close()
        self._dev = None

    MAX_CALLS_IN_FLIGHT = 3  # retrieve() calls of different threads that may be on the GPU at the same time

    @contextlib.contextmanager
    def _lane(self, dev, base_only: bool):
        """A handle for one retrieve call. Calls from several threads (a server) overlap on the GPU: the first takes the
        index's own handle, the next ones views of it (same resident arrays, own stream and workspaces), at most
        MAX_CALLS_IN_FLIGHT in all; further callers wait. A call that involves a weight mask needs the handle the mask
        lives on (`base_only`); several devices (DeviceGroup) have one handle."""
        if not isinstance(dev, _lib.DeviceIndex):
            yield dev
            return
        with self._lanes_cv:
            st = self._lanes
            if st is None or st["dev"] is not dev:
                st = self._lanes = {"dev": dev, "free": [dev], "views": []}
            while True:
                if base_only:
                    if dev in st["free"]:
                        st["free"].remove(dev)
                        h = dev
                        break
                elif st["free"]:
                    h = st["free"].pop()  # the index's own handle sits at the end: one caller at a time never makes a view
                    break
                elif len(st["views"]) + 1 < self.MAX_CALLS_IN_FLIGHT:
                    h = dev.view()
                    st["views"].append(h)
                    break
                self._lanes_cv.wait()
            busy = 1 + len(st["views"]) - len(st["free"])  # handles in use, this one included
        try:
            # the library sizes a batch's parts by the batches in flight; by default that is the number of handles, which
            # stays up after a burst of callers - tell it how many calls are on the GPU right now
            if getattr(h, "_in_flight_told", None) != busy:
                h.set_option("IN_FLIGHT", busy)
                h._in_flight_told = busy
            yield h
        finally:
            with self._lanes_cv:
                if self._lanes is st:
                    if h is dev:
                        st["free"].append(h)
                    else:
                        st["free"].insert(0, h)
                    self._lanes_cv.notify_all()

    # ------------------------------------------------------------------ weight masks (bm25s/__init__.py:610-612, 833-845)
    def _check_mask(self, weight_mask):
        if not isinstance(weight_mask, np.ndarray):
            raise ValueError("weight_mask must be a numpy array.")
        if weight_mask.ndim != 1:
            raise ValueError("weight_mask must be a 1D array.")
        if len(weight_mask) != self.scores["num_docs"]:
            raise ValueError("The length of the weight_mask must be the same as the length of the corpus.")

    def _load_mask(self, dev, weight_mask, sticky):
        """Make `weight_mask` the device-resident mask of `dev` (bit-packed if it is a 0/1 mask). A mask that cannot have
        changed since it was loaded - the same read-only array object - is not packed or uploaded again."""
        key = (id(weight_mask), weight_mask.__array_interface__["data"][0], weight_mask.dtype.str, len(weight_mask), id(dev))
        if self._mask_key == key and not weight_mask.flags.writeable:
            self._mask_sticky = sticky
            return
        kind, arr = _lib.pack_weight_mask(weight_mask)
        dev.set_mask(kind, arr)
        self._mask_key = key
        self._mask_sticky = sticky
        self._mask_keepalive = weight_mask  # keeps the id from being reused while it is the key

    def set_weight_mask(self, weight_mask):
        """Extension: keep `weight_mask` resident on the device; every later `retrieve` without a `weight_mask` argument is
        filtered / weighted by it (no per-call packing or upload). `clear_weight_mask()` removes it."""
        self._check_mask(weight_mask)
        with self._mask_lock():
            with self._lane(self._device_index(), base_only=True) as h:  # no call is running on the handle meanwhile
                self._load_mask(h, weight_mask, sticky=True)

    def clear_weight_mask(self):
        with self._mask_lock():
            if self._dev is not None and self._mask_key is not None:
                with self._lane(self._dev, base_only=True) as h:
                    h.set_mask(_lib.MASK_NONE)
            self._mask_key, self._mask_sticky, self._mask_keepalive = None, False, None

    def _mask_lock(self):
        lk = getattr(self, "_mask_lk", None)
        if lk is None:
            lk = self._mask_lk = threading.RLock()
        return lk

    def _device_index(self) -> _lib.DeviceIndex:
        """Upload `self.scores` on first use; re-upload if the arrays were swapped (load_scores, direct assignment)."""
        with self._dev_lock:  # concurrent retrieve() calls: one upload
            key = self._residency_key()
            if self._dev is None or not self._same_key(self._dev_key, key):
                self._close_dev()
                sc = self.scores
                if self.devices is not None and len(self.devices) > 1:
                    self._dev = _lib.DeviceGroup.upload(sc["data"], sc["indices"], sc["indptr"], int(sc["num_docs"]),
                                                        nonocc=self.nonoccurrence_array, devices=self.devices)
                else:
                    self._dev = _lib.DeviceIndex.upload(sc["data"], sc["indices"], sc["indptr"], int(sc["num_docs"]),
                                                        nonocc=self.nonoccurrence_array, device=self.device)
                self._dev_key = key
                if self._mask_sticky and self._mask_keepalive is not None:  # a new device copy: the resident mask moves along
                    kind, arr = _lib.pack_weight_mask(self._mask_keepalive)
                    self._dev.set_mask(kind, arr)
                    mk = self._mask_keepalive
                    self._mask_key = (id(mk), mk.__array_interface__["data"][0], mk.dtype.str, len(mk), id(self._dev))
                else:
                    self._mask_key = None
            return self._dev

    def _adopt_device_index(self, dev: _lib.DeviceIndex, num_docs: int):
        """After a GPU build: mirror the arrays to host so `self.scores` / save() are reference-identical."""
        data, indices, indptr, nonocc = dev.download()
        with self._dev_lock:
            if self._dev is not None and self._dev is not dev:
                self._close_dev()
            self.scores = {"data": data, "indices": indices, "indptr": indptr, "num_docs": num_docs}
            self.nonoccurrence_array = nonocc
            self._dev = dev
            self._dev_key = self._residency_key()

    # ------------------------------------------------------------------ index build
    @staticmethod
    def _infer_corpus_object(corpus):
        """bm25s/__init__.py:246-270"""
        msg = ("Corpus must be a list of list of tokens, an object with the `ids` and `vocab` attributes, or a tuple of "
               "two lists: the first list is the list of unique token IDs, and the second list is the list of token "
               "IDs for each document.")
        if hasattr(corpus, "ids") and hasattr(corpus, "vocab"):
            return "object"
        elif isinstance(corpus, tuple) and len(corpus) == 2:
            c1, c2 = corpus
            if isinstance(c1, list) and isinstance(c2, dict):
                return "tuple"
            raise ValueError(msg)
        elif isinstance(corpus, Iterable):
            return "token_ids" if is_list_of_list_of_type(corpus, type_=int) else "tokens"
        raise ValueError(msg)

    def build_index_from_ids(self, unique_token_ids: List[int], corpus_token_ids: List[List[int]],
                             show_progress=True, leave_progress=False):
        """GPU replacement of bm25s/__init__.py:326-438: df, idf, tfc, score triples and CSC assembly run in
        libbm25cu (bm25cu_index_build). Returns the reference's `scores` dict (host mirrors of the HBM arrays)."""
        if self.method not in _METHODS:
            raise ValueError(f"Invalid score_tfc value: {self.method}. Choose from 'robertson', 'lucene', 'atire'.")
        if self.idf_method not in _METHODS:
            raise ValueError(
                f"Invalid score_idf_inner value: {self.idf_method}. Choose from 'robertson', 'lucene', 'atire', 'bm25l', 'bm25+'."
            )
        n_vocab = len(unique_token_ids)
        tokens, doc_ptr = _flatten(corpus_token_ids)
        if len(tokens) and (tokens.min() < 0 or tokens.max() >= n_vocab):
            raise ValueError("corpus token ids must lie in [0, len(unique_token_ids))")
        if self.devices is not None and len(self.devices) > 1:
            dev = _lib.DeviceGroup.build(tokens.astype(np.int32), doc_ptr, n_vocab, self.method, self.idf_method,
                                         float(self.k1), float(self.b), float(self.delta), devices=self.devices)
        else:
            dev = _lib.DeviceIndex.build(tokens.astype(np.int32), doc_ptr, n_vocab, self.method, self.idf_method,
                                         float(self.k1), float(self.b), float(self.delta), device=self.device)
        self._adopt_device_index(dev, len(corpus_token_ids))
        return self.scores

    def build_index_from_tokens(self, corpus_tokens, show_progress=True, leave_progress=False):
        """bm25s/__init__.py:440-473 (vocabulary numbered in first-seen order here, which is reproducible)."""
        vocab_dict: Dict[str, int] = {}
        corpus_token_ids = []
        for doc in corpus_tokens:
            ids = []
            for tok in doc:
                i = vocab_dict.get(tok)
                if i is None:
                    i = vocab_dict[tok] = len(vocab_dict)
                ids.append(i)
            corpus_token_ids.append(ids)
        scores = self.build_index_from_ids(list(vocab_dict.values()), corpus_token_ids, show_progress, leave_progress)
        return scores, vocab_dict

    def index(self, corpus, create_empty_token=True, show_progress=True, leave_progress=False):
        """bm25s/__init__.py:475-570"""
        inferred = self._infer_corpus_object(corpus)
        if inferred == "tokens":
            scores, vocab_dict = self.build_index_from_tokens(corpus, show_progress, leave_progress)
        else:
            if inferred == "tuple":
                corpus_token_ids, vocab_dict = corpus
            elif inferred == "object":
                corpus_token_ids, vocab_dict = corpus.ids, corpus.vocab
            else:  # "token_ids": ids are used as columns as they are (:535-547)
                corpus_token_ids = corpus
                unique_ids = set()
                for doc_ids in corpus_token_ids:
                    unique_ids.update(doc_ids)
                if create_empty_token:
                    if 0 not in unique_ids:
                        unique_ids.add(0)
                    else:
                        unique_ids.add(max(unique_ids) + 1)
                vocab_dict = {token_id: i for i, token_id in enumerate(unique_ids)}
            unique_token_ids = list(vocab_dict.values())
            scores = self.build_index_from_ids(unique_token_ids, corpus_token_ids, show_progress, leave_progress)
        if create_empty_token:
            if inferred != "token_ids" and "" not in vocab_dict:
                vocab_dict[""] = max(vocab_dict.values()) + 1
        self.scores = scores
        self.vocab_dict = vocab_dict
        self.unique_token_ids_set = set(self.vocab_dict.values())

    # ------------------------------------------------------------------ scoring
    def get_tokens_ids(self, query_tokens: List[str]) -> List[int]:
        """bm25s/__init__.py:572-579 (OOV tokens dropped)"""
        vd = self.vocab_dict
        return [vd[token] for token in query_tokens if token in vd]

    _EMPTY_QUERY_MSG = (
        "The query does not contain any tokens that are in the vocabulary. "
        "Please provide a query that contains at least one token that is in the vocabulary. "
        "Alternatively, you can set `create_empty_token=True` when calling `index` to add an empty token to the vocabulary. "
        "You can also manually add an empty token to the vocabulary by setting `retriever.vocab_dict[''] = max(retriever.vocab_dict.values()) + 1`. "
        "Then, run `retriever.unique_token_ids_set = set(retriever.vocab_dict.values())` to update the unique token IDs."
    )

    def _membership_table(self) -> np.ndarray:
        """Boolean table over token ids: True where the id is in `unique_token_ids_set` (cached per set object)."""
        uts = self.unique_token_ids_set
        key = (id(uts), len(uts))
        cached = getattr(self, "_member_cache", None)
        if cached is None or cached[0] != key:
            ids = np.fromiter((t for t in uts if isinstance(t, (int, np.integer)) and t >= 0), dtype=np.int64)
            table = np.zeros(int(ids.max()) + 1 if len(ids) else 0, dtype=bool)
            table[ids] = True
            self._member_cache = cached = (key, table)
        return cached[1]

    def _filter_int_queries(self, queries) -> Tuple[np.ndarray, np.ndarray]:
        """list[list[int]] -> (flat, pointers) keeping ids of `unique_token_ids_set`; a query left empty becomes the
        empty-token query, or raises as the reference does (bm25s/__init__.py:778-798)."""
        flat, ptr = _flatten(queries)
        member = self._membership_table()
        keep = (flat >= 0) & (flat < len(member))
        keep[keep] = member[flat[keep]]
        csum = np.zeros(len(flat) + 1, dtype=np.int64)
        np.cumsum(keep, out=csum[1:])
        flat, ptr = flat[keep], csum[ptr]
        empty = np.nonzero(np.diff(ptr) == 0)[0]
        if len(empty):
            if "" not in self.vocab_dict:
                raise ValueError(self._EMPTY_QUERY_MSG)
            flat = np.insert(flat, ptr[empty], self.vocab_dict[""])
            add = np.zeros(len(ptr), dtype=np.int64)
            add[empty + 1] = 1
            ptr = ptr + np.cumsum(add)
        return flat, ptr

    def _remap_tokenized(self, tokenized: Tokenized) -> Tuple[np.ndarray, np.ndarray]:
        """Tokenized(ids, vocab) -> (flat index ids, pointers) with out-of-vocabulary tokens dropped."""
        vocab = tokenized.vocab
        size = (max(vocab.values()) + 1) if len(vocab) else 0
        remap = np.full(size, -2, dtype=np.int64)  # -2: id absent from the query vocabulary, -1: OOV for the index
        get = self.vocab_dict.get
        for token, qid in vocab.items():
            remap[qid] = get(token, -1)
        flat_q, ptr = _flatten(tokenized.ids)
        if len(flat_q) == 0:
            return flat_q.astype(np.int32), ptr
        if int(flat_q.min()) < 0 or int(flat_q.max()) >= size:
            bad = flat_q[(flat_q < 0) | (flat_q >= size)][0]
            raise KeyError(int(bad))  # the reference fails in reverse_vocab[token_id]
        mapped = remap[flat_q]
        if (mapped == -2).any():
            raise KeyError(int(flat_q[mapped == -2][0]))
        keep = mapped >= 0
        csum = np.zeros(len(mapped) + 1, dtype=np.int64)
        np.cumsum(keep, out=csum[1:])
        return mapped[keep], csum[ptr]

    def _check_max_token_id(self, ids: np.ndarray):
        max_token_id = int(ids.max(initial=0))
        if max_token_id >= len(self.scores["indptr"]) - 1:
            raise ValueError(
                f"The maximum token ID in the query ({max_token_id}) is higher than the number of tokens in the index."
                "This likely means that the query contains tokens that are not in the index."
            )

    def get_scores_from_ids(self, query_tokens_ids: List[int], weight_mask=None) -> np.ndarray:
        """bm25s/__init__.py:581-620 on the device (bm25cu_scores_dense): float32[num_docs]."""
        ids = np.asarray(query_tokens_ids, dtype=np.int32)
        self._check_max_token_id(ids)
        shift = None
        if self.nonoccurrence_array is not None:
            shift = self.nonoccurrence_array[ids].sum()  # numpy's own float32 reduction (bit-exact by construction)
        mask = self._mask_f64(weight_mask)
        return self._device_index().scores_dense(ids, shift=shift, mask=mask)

    def _mask_f64(self, weight_mask):
        """The weight mask as the float64 vector the C ABI takes (exact for bool / int32 / float32 / float64 masks); its
        length is checked here because the library reads num_docs values from it."""
        if weight_mask is None:
            return None
        mask = np.asarray(weight_mask)
        if mask.ndim != 1:
            raise ValueError("weight_mask must be a 1D array.")
        if len(mask) != self.scores["num_docs"]:
            raise ValueError("The length of the weight_mask must be the same as the length of the corpus.")
        return np.ascontiguousarray(mask, dtype=np.float64)

    def get_scores(self, query_tokens_single: List[str], weight_mask=None) -> np.ndarray:
        """bm25s/__init__.py:622-638"""
        if not isinstance(query_tokens_single, list):
            raise ValueError("The query_tokens must be a list of tokens.")
        if isinstance(query_tokens_single[0], str):
            query_tokens_ids = self.get_tokens_ids(query_tokens_single)
        elif isinstance(query_tokens_single[0], (int, np.integer)):
            query_tokens_ids = query_tokens_single
        else:
            raise ValueError("The query_tokens must be a list of tokens or a list of token IDs.")
        return self.get_scores_from_ids(query_tokens_ids, weight_mask=weight_mask)

    def _query_shifts(self, flat: np.ndarray, ptr: np.ndarray):
        """Per-query `nonoccurrence_array[ids].sum()` (bm25s/__init__.py:616-617), computed by numpy itself, one
        2-D reduction per distinct query length (same pairwise routine per row as the 1-D sum)."""
        if self.nonoccurrence_array is None:
            return None
        lens = np.diff(ptr)
        shift = np.zeros(len(lens), dtype=np.float32)
        vals = self.nonoccurrence_array[flat] if len(flat) else np.zeros(0, np.float32)
        for L in np.unique(lens):
            if L == 0:
                continue
            rows = np.nonzero(lens == L)[0]
            gather = ptr[rows][:, None] + np.arange(L)[None, :]
            shift[rows] = np.ascontiguousarray(vals[gather]).sum(axis=1)
        return shift

    def retrieve(self, query_tokens, corpus: List[Any] = None, k: int = 10, sorted: bool = True,
                 return_as: str = "tuple", show_progress: bool = True, leave_progress: bool = False,
                 n_threads: int = 0, chunksize: int = 50, backend_selection: str = "auto",
                 weight_mask: np.ndarray = None):
        """Top-k documents per query: bm25s/__init__.py:676-939 with the per-query scorer + selection replaced by
        one batched call into libbm25cu (bm25cu_retrieve). `n_threads`, `chunksize`, `backend_selection` and the
        progress flags are accepted and ignored."""
        num_docs = self.scores["num_docs"]
        if k > num_docs:
            raise ValueError(
                f"k of {k} is larger than the number of available scores"
                f", which is {num_docs} (corpus size should be larger than top-k)."
                f" Please set with a smaller k or increase the size of corpus."
            )
        if return_as not in ["tuple", "documents"]:
            raise ValueError("`return_as` must be either 'tuple' or 'documents'")

        int_flat = None
        if is_list_of_list_of_type(query_tokens, type_=int):
            if not hasattr(self, "unique_token_ids_set") or self.unique_token_ids_set is None:
                raise ValueError(
                    "The unique_token_ids_set attribute is not found. Please run the `index` method first, or make sure"
                    "run retriever.load(load_vocab=True) before calling retrieve on list of list of token IDs (int)."
                )
            # reference :778-798 filters token by token against the set; here the batch is flattened first and
            # filtered through a boolean membership table (same result, one numpy gather)
            int_flat = self._filter_int_queries(query_tokens)

        if isinstance(query_tokens, tuple) and not _is_tuple_of_list_of_tokens(query_tokens):
            if len(query_tokens) != 2:
                raise ValueError(
                    "Expected a list of string or a tuple of two elements: the first element is the "
                    "list of unique token IDs, "
                    "and the second element is the list of token IDs for each document."
                    f"Found {len(query_tokens)} elements instead."
                )
            ids, vocab = query_tokens
            if not isinstance(ids, Iterable):
                raise ValueError("The first element of the tuple passed to retrieve must be an iterable.")
            if not isinstance(vocab, dict):
                raise ValueError("The second element of the tuple passed to retrieve must be a dictionary.")
            query_tokens = Tokenized(ids=ids, vocab=vocab)

        tokenized_flat = None
        if isinstance(query_tokens, Tokenized):
            # reference: ids -> strings -> ids, token by token (tokenization.py:513-521, __init__.py:572-579, 857-860);
            # here: one remap table per call (query-vocabulary id -> index id, -1 = out of vocabulary), applied to the
            # flattened ids by numpy. Same result: OOV tokens are dropped, order and duplicates are kept.
            tokenized_flat = self._remap_tokenized(query_tokens)

        corpus = corpus if corpus is not None else self.corpus

        if weight_mask is not None:
            if not isinstance(weight_mask, np.ndarray):
                raise ValueError("weight_mask must be a numpy array.")
            if weight_mask.ndim != 1:
                raise ValueError("weight_mask must be a 1D array.")
            if len(weight_mask) != self.scores["num_docs"]:
                raise ValueError("The length of the weight_mask must be the same as the length of the corpus.")

        # ---- the cuda branch sits where the numba branch is in the reference (:847-887)
        if tokenized_flat is not None:
            flat, ptr = tokenized_flat
        elif int_flat is not None:
            flat, ptr = int_flat
        elif isinstance(query_tokens, (list, tuple)) and all(isinstance(q, (list, tuple)) for q in query_tokens):
            query_tokens_ids = [
                self.get_tokens_ids(q) if (len(q) and isinstance(q[0], str)) else list(q) for q in query_tokens
            ]
        else:
            raise ValueError(
                "The query_tokens must be a list of list of tokens (str for stemmed words, int for token ids matching corpus) or a tuple of two lists: the first list is the list of unique token IDs, and the second list is the list of token IDs for each document."
            )
        if tokenized_flat is None and int_flat is None:
            flat, ptr = _flatten(query_tokens_ids)
        scores, indices = self.retrieve_flat(flat, ptr, k=k, weight_mask=weight_mask)

        if corpus is None:
            retrieved_docs = indices
        else:
            if isinstance(corpus, JsonlCorpus):
                retrieved_docs = corpus[indices]
            elif isinstance(corpus, np.ndarray) and corpus.ndim == 1:
                retrieved_docs = corpus[indices]
            else:
                index_flat = indices.flatten().tolist()
                results = [corpus[i] for i in index_flat]
                retrieved_docs = np.array(results).reshape(indices.shape)

        if return_as == "tuple":
            return Results(documents=retrieved_docs, scores=scores)
        return retrieved_docs

    def retrieve_flat(self, flat_ids: np.ndarray, ptr: np.ndarray, k: int = 10, weight_mask=None, with_stats=False):
        """Array entry point (no Python lists): flat token ids + int64 pointers -> (scores f32[Q,k], ids i64[Q,k])."""
        flat = np.asarray(flat_ids)
        ptr = np.asarray(ptr, dtype=np.int64)
        nq = len(ptr) - 1
        if len(flat):
            if flat.dtype.kind not in "iu":
                raise ValueError("token ids must be integers")
            if int(flat.min()) < 0:
                raise ValueError("token ids must be non-negative")
            self._check_max_token_id(flat)
        flat = flat.astype(np.int32, copy=False)
        if k == 0 or nq == 0:
            out = (np.zeros((nq, k), np.float32), np.zeros((nq, k), np.int64))
            return out + ({},) if with_stats else out
        shift = self._query_shifts(flat, ptr)
        dev = self._device_index()
        with contextlib.ExitStack() as stack:
            with self._mask_lock():
                if weight_mask is not None:
                    self._check_mask(weight_mask)
                elif self._mask_key is not None and not self._mask_sticky:
                    with self._lane(dev, base_only=True) as h:
                        h.set_mask(_lib.MASK_NONE)  # the mask of an earlier call does not outlive it
                    self._mask_key, self._mask_keepalive = None, None
                masked = weight_mask is not None or self._mask_key is not None
                # the handle is taken under the mask lock: while an unmasked call holds the index's own handle no other
                # call can make a mask resident on it
                h = stack.enter_context(self._lane(dev, base_only=masked))
                if masked:
                    # masked calls run one at a time, on the handle the mask is resident on (1 bit per document for 0/1 masks)
                    if weight_mask is not None:
                        self._load_mask(h, weight_mask, sticky=False)
                    res = h.retrieve(flat, ptr, k, shift=shift, mask=None, with_stats=with_stats)
            if not masked:  # no mask in force: concurrent callers overlap, each on a handle of its own
                res = h.retrieve(flat, ptr, k, shift=shift, mask=None, with_stats=with_stats)
        ids, scores = res[0], res[1]
        out = (scores, ids.astype(np.int64))
        return out + (res[2],) if with_stats else out

    # ------------------------------------------------------------------ persistence (bm25s/__init__.py:941-1282)
    def save(self, save_dir, corpus=None, data_name="data.csc.index.npy", indices_name="indices.csc.index.npy",
             indptr_name="indptr.csc.index.npy", vocab_name="vocab.index.json", params_name="params.index.json",
             nnoc_name="nonoccurrence_array.index.npy", corpus_name="corpus.jsonl", allow_pickle=False,
             show_progress=True, leave_progress=False):
        save_dir = Path(save_dir)
        save_dir.mkdir(parents=True, exist_ok=True)
        np.save(save_dir / data_name, self.scores["data"], allow_pickle=allow_pickle)
        np.save(save_dir / indices_name, self.scores["indices"], allow_pickle=allow_pickle)
        np.save(save_dir / indptr_name, self.scores["indptr"], allow_pickle=allow_pickle)
        if self.nonoccurrence_array is not None:
            np.save(save_dir / nnoc_name, self.nonoccurrence_array, allow_pickle=allow_pickle)
        with open(save_dir / vocab_name, "wt", encoding="utf-8") as f:
            f.write(json.dumps(self.vocab_dict, ensure_ascii=False))
        params = dict(k1=self.k1, b=self.b, delta=self.delta, method=self.method, idf_method=self.idf_method,
                      dtype=self.dtype, int_dtype=self.int_dtype, num_docs=self.scores["num_docs"],
                      version=__version__, backend=self.backend)
        with open(save_dir / params_name, "w") as f:
            json.dump(params, f, indent=4)
        corpus = corpus if corpus is not None else self.corpus
        if corpus is not None:
            with open(save_dir / corpus_name, "wt", encoding="utf-8") as f:
                if not isinstance(corpus, Iterable):
                    logging.warning("The corpus is not an iterable. Skipping saving the corpus.")
                for i, doc in enumerate(corpus):
                    if isinstance(doc, str):
                        doc = {"id": i, "text": doc}
                    elif isinstance(doc, (dict, list, tuple)):
                        pass
                    else:
                        logging.warning(f"Document at index {i} is not a string, dictionary, list or tuple. Skipping.")
                        continue
                    try:
                        doc_str = json.dumps(doc, ensure_ascii=False)
                    except Exception as e:  # noqa: BLE001
                        logging.warning(f"Error saving document at index {i}: {e}")
                    else:
                        f.write(doc_str + "\n")
            mmidx = find_newline_positions(save_dir / corpus_name)
            save_mmindex(mmidx, path=save_dir / corpus_name)

    def load_scores(self, save_dir, data_name="data.csc.index.npy", indices_name="indices.csc.index.npy",
                    indptr_name="indptr.csc.index.npy", num_docs=None, mmap=False, allow_pickle=False):
        """bm25s/__init__.py:1074-1127; the upload to HBM happens lazily on the next retrieve/get_scores."""
        save_dir = Path(save_dir)
        mmap_mode = "r" if mmap else None
        scores = {
            "data": np.load(save_dir / data_name, allow_pickle=allow_pickle, mmap_mode=mmap_mode),
            "indices": np.load(save_dir / indices_name, allow_pickle=allow_pickle, mmap_mode=mmap_mode),
            "indptr": np.load(save_dir / indptr_name, allow_pickle=allow_pickle, mmap_mode=mmap_mode),
            "num_docs": num_docs,
        }
        self.scores = scores
        self.invalidate()

    @classmethod
    def load(cls, save_dir, data_name="data.csc.index.npy", indices_name="indices.csc.index.npy",
             indptr_name="indptr.csc.index.npy", vocab_name="vocab.index.json", params_name="params.index.json",
             nnoc_name="nonoccurrence_array.index.npy", corpus_name="corpus.jsonl", load_corpus=False, mmap=False,
             allow_pickle=False, load_vocab=True, override_params: dict = None, show_progress=True,
             leave_progress=False, **kwargs):
        """bm25s/__init__.py:1129-1282. An index saved by the reference (backend "numpy"/"numba" in its params)
        loads here too: any non-cuda backend recorded in the file is replaced by "cuda"."""
        if not isinstance(mmap, bool):
            raise ValueError("`mmap` must be a boolean")
        save_dir = Path(save_dir)
        with open(save_dir / params_name, "r") as f:
            params: dict = json.loads(f.read())
        if override_params is not None:
            params.update(override_params)
        if kwargs:
            params.update(kwargs)
        if load_vocab:
            with open(save_dir / vocab_name, "r", encoding="utf-8") as f:
                vocab_dict: dict = json.loads(f.read())
        else:
            vocab_dict = {}
        original_version = params.pop("version", None)
        num_docs = params.pop("num_docs", None)
        if params.get("backend", "cuda") != "cuda":
            params["backend"] = "cuda"
        bm25_obj = cls(**params)
        bm25_obj.vocab_dict = vocab_dict
        bm25_obj._original_version = original_version
        bm25_obj.unique_token_ids_set = set(bm25_obj.vocab_dict.values())
        bm25_obj.load_scores(save_dir=save_dir, data_name=data_name, indices_name=indices_name,
                             indptr_name=indptr_name, mmap=mmap, num_docs=num_docs, allow_pickle=allow_pickle)
        if load_corpus:
            corpus_file = save_dir / corpus_name
            if os.path.exists(corpus_file):
                if mmap is True:
                    corpus = JsonlCorpus(corpus_file)
                else:
                    corpus = []
                    with open(corpus_file, "r", encoding="utf-8") as f:
                        for line in f:
                            corpus.append(json.loads(line))
                bm25_obj.corpus = corpus
        if bm25_obj.method in bm25_obj.methods_requiring_nonoccurrence:
            nnm_path = save_dir / nnoc_name
            if nnm_path.exists():
                bm25_obj.nonoccurrence_array = np.load(nnm_path, allow_pickle=allow_pickle)
            else:
                raise FileNotFoundError(f"Non-occurrence array not found at {nnm_path}")
        else:
            bm25_obj.nonoccurrence_array = None
        return bm25_obj
