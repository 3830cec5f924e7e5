"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONPATH=/root/reference DISABLE_TQDM=1 python oracle/make_golden.py

The fixtures are small, committed, and are what pins the oracle (oracle/oracle.py, oracle/bm25_oracle.c)
and the CUDA path when the reference is absent. Nothing at test/bench run time reads /root/reference.
"""
import os
import sys

os.environ.setdefault("DISABLE_TQDM", "1")
REF = os.environ.get("BM25S_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bm25s  # noqa: E402  (the reference)
from bm25s_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
METHODS = ["robertson", "lucene", "atire", "bm25l", "bm25+"]


def flat(lists):
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in lists], out=ptr[1:])
    tok = np.concatenate([np.asarray(x, dtype=np.int32) for x in lists]) if ptr[-1] else np.zeros(0, np.int32)
    return tok.astype(np.int32), ptr


def index_ref(corpus_lists, n_vocab, method, **kw):
    r = bm25s.BM25(method=method, **kw)
    vocab = {i: i for i in range(n_vocab)}
    r.index((corpus_lists, vocab), create_empty_token=False, show_progress=False)
    return r


def readme_fixture():
    """the README corpus, 5 methods, dense scores + CSC + retrieve."""
    corpus = [
        "a cat is a feline and likes to purr",
        "a dog is the human's best friend and loves to play",
        "a bird is a beautiful animal that can fly",
        "a fish is a creature that lives in water and swims",
    ]
    tok = bm25s.tokenize(corpus, stopwords="en", show_progress=False)
    ids, vocab = tok.ids, tok.vocab
    q = bm25s.tokenize(["cat feline dog bird fish"], stopwords="en", show_progress=False)
    rev = {v: k for k, v in q.vocab.items()}
    q_words = [rev[i] for i in q.ids[0]]
    q_ids = [vocab[w] for w in q_words]
    out = {"corpus_tokens": flat(ids)[0], "corpus_ptr": flat(ids)[1], "query_ids": np.asarray(q_ids, np.int32),
           "n_vocab": np.int64(len(vocab))}
    for m in METHODS:
        r = bm25s.BM25(method=m)
        r.index((ids, dict(vocab)), create_empty_token=False, show_progress=False)
        key = m.replace("+", "plus")
        out[f"{key}_data"] = r.scores["data"]
        out[f"{key}_indices"] = r.scores["indices"]
        out[f"{key}_indptr"] = r.scores["indptr"]
        if r.nonoccurrence_array is not None:
            out[f"{key}_nonocc"] = r.nonoccurrence_array
        out[f"{key}_dense"] = r.get_scores(q_ids)
        for mk, mask in (("f32", np.array([1, 0, 0, 1], np.float32)), ("i32", np.array([1, 0, 0, 1], np.int32)),
                         ("bool", np.array([1, 0, 0, 1], np.bool_))):
            out[f"{key}_dense_mask_{mk}"] = r.get_scores(q_ids, weight_mask=mask)
        res = r.retrieve([q_ids], k=2, show_progress=False)
        out[f"{key}_top2_ids"] = res.documents
        out[f"{key}_top2_scores"] = res.scores
    np.savez_compressed(os.path.join(OUT, "readme_corpus.npz"), **out)


def synth_fixture(name, n_docs, avg_len, n_vocab, n_queries, k, seed, dense_queries=8):
    corp = synth.make_corpus(n_docs, avg_len, n_vocab, seed=seed)
    qs = synth.make_queries(n_queries, n_vocab, seed=seed + 1)
    corpus_lists, query_lists = corp.to_lists(), qs.to_lists()
    rng = np.random.default_rng(seed + 2)
    mask01 = (rng.random(n_docs) < 0.7).astype(np.float32)
    maskf = rng.random(n_docs).astype(np.float64)
    out = {"corpus_tokens": corp.tokens, "corpus_ptr": corp.ptr, "query_tokens": qs.tokens, "query_ptr": qs.ptr,
           "n_vocab": np.int64(n_vocab), "k": np.int64(k), "mask01": mask01, "maskf64": maskf}
    for m in METHODS:
        key = m.replace("+", "plus")
        r = index_ref(corpus_lists, n_vocab, m)
        out[f"{key}_data"] = r.scores["data"]
        out[f"{key}_indices"] = r.scores["indices"].astype(np.int32)
        out[f"{key}_indptr"] = r.scores["indptr"].astype(np.int64)
        if r.nonoccurrence_array is not None:
            out[f"{key}_nonocc"] = r.nonoccurrence_array
        res = r.retrieve(query_lists, k=k, show_progress=False)
        out[f"{key}_ids"] = res.documents.astype(np.int64)
        out[f"{key}_scores"] = res.scores
        res = r.retrieve(query_lists, k=k, show_progress=False, weight_mask=mask01)
        out[f"{key}_ids_mask01"] = res.documents.astype(np.int64)
        out[f"{key}_scores_mask01"] = res.scores
        res = r.retrieve(query_lists, k=k, show_progress=False, weight_mask=maskf)
        out[f"{key}_scores_maskf64"] = res.scores
        dense = np.stack([r.get_scores(q) for q in query_lists[:dense_queries]])
        out[f"{key}_dense"] = dense
        out[f"{key}_dense_maskf64"] = np.stack([r.get_scores(q, weight_mask=maskf) for q in query_lists[:dense_queries]])
    # non-default hyper-parameters exercise the float32 numerators (k1 + 1 not representable)
    r = bm25s.BM25(method="atire", k1=1.2, b=0.6)
    r.index((corpus_lists, {i: i for i in range(n_vocab)}), create_empty_token=False, show_progress=False)
    out["atire_k12_data"] = r.scores["data"]
    r = bm25s.BM25(method="bm25+", k1=0.9, b=0.4, delta=1.0)
    r.index((corpus_lists, {i: i for i in range(n_vocab)}), create_empty_token=False, show_progress=False)
    out["bm25plus_k09_data"] = r.scores["data"]
    out["bm25plus_k09_nonocc"] = r.nonoccurrence_array
    r = bm25s.BM25(method="lucene", idf_method="robertson")
    r.index((corpus_lists, {i: i for i in range(n_vocab)}), create_empty_token=False, show_progress=False)
    out["lucene_idfrobertson_data"] = r.scores["data"]
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)


def known_answer_fixture():
    """The reference's own unit-test vectors, re-derived by running the reference functions."""
    from bm25s.scoring import _compute_relevance_from_scores_jit_ready, _compute_relevance_from_scores_legacy
    from bm25s.selection import topk

    out = {}
    # tests/core/test_scoring.py:26-98
    data = np.array([1.0, 2.0], np.float32)
    indices = np.array([0, 1], np.int32)
    indptr = np.array([0, 1, 2], np.int32)
    out["ka1_legacy"] = _compute_relevance_from_scores_legacy(data, indptr, indices, 3, [0, 1], np.float32)
    out["ka1_jit"] = _compute_relevance_from_scores_jit_ready(data, indptr, indices, 3, np.array([0, 1], np.int32), np.float32)
    out["ka1_empty"] = _compute_relevance_from_scores_legacy(data, indptr, indices, 3, [], np.float32)
    data2 = np.array([1.5, 2.5], np.float32)
    indices2 = np.array([0, 0], np.int32)
    out["ka2_jit"] = _compute_relevance_from_scores_jit_ready(data2, indptr, indices2, 2, np.array([0, 1], np.int32), np.float32)
    # tests/core/test_topk.py:16-59
    np.random.seed(42)
    scores = np.random.uniform(-10, 10, 2000)
    s, i = topk(scores, 5, backend="numpy", sorted=True)
    out["topk_scores_in"] = scores
    out["topk_s"] = s
    out["topk_i"] = i
    # tests/core/test_selection.py:34-58
    s, i = topk(np.array([3.0, 1.0, 4.0, 2.0], np.float32), k=2, backend="numpy", sorted=True)
    out["sel_s"] = s
    out["sel_i"] = i
    np.savez_compressed(os.path.join(OUT, "known_answers.npz"), **out)


if __name__ == "__main__":
    readme_fixture()
    known_answer_fixture()
    synth_fixture("synth_small", n_docs=300, avg_len=20, n_vocab=500, n_queries=40, k=10, seed=10)
    synth_fixture("synth_medium", n_docs=3000, avg_len=40, n_vocab=4000, n_queries=60, k=25, seed=20)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
