"""CPU simulation (numpy, no GPU) of the work K1m does per query - postings streamed, survivors looked up - under
the current bounds (one upper bound per column: its maximum) and under impact tiers (the
postings of a long column above a score quantile form a HIGH tier that is scored first; the rest, the LOW tier, is then
bounded by the quantile instead of the maximum). Both variants are exact; the simulation checks its top-k against a
brute-force sum. Thresholds move once per chunk of 2048 streamed postings, as in the kernel (one part per query).

usage: python tools/tier_sim.py [n_docs] [n_queries] [k]      (defaults 880000 1000 10; V = 2 M, Poisson(56) lengths)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bm25s_b200 import synth  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 880_000
n_queries = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
V, L = 2_000_000, 56
TIER_MIN_DF, TIER_FRAC = int(os.environ.get("TIER_MIN_DF", 2048)), 1.0 / float(os.environ.get("TIER_INV_FRAC", 32))
CHUNK = 2048

t0 = time.time()
corp = synth.make_corpus(n_docs, L, V, seed=0)
doc_of = np.repeat(np.arange(n_docs, dtype=np.int32), np.diff(corp.ptr))
key = corp.tokens.astype(np.int64) * n_docs + doc_of
key.sort()
uniq, tf = np.unique(key, return_counts=True)
tok = (uniq // n_docs).astype(np.int32)
doc = (uniq % n_docs).astype(np.int32)
indptr = np.zeros(V + 1, np.int64)
np.cumsum(np.bincount(tok, minlength=V), out=indptr[1:])
df = np.diff(indptr)
dl = np.diff(corp.ptr).astype(np.float64)
idf = np.log(1.0 + (n_docs - df + 0.5) / (df + 0.5))
data = (idf[tok] * tf / (tf + 1.5 * (0.25 + 0.75 * dl[doc] / dl.mean()))).astype(np.float32)  # lucene variant
del key, uniq, doc_of
qs = synth.make_queries(n_queries, V, seed=1)
print(f"corpus {n_docs} docs, {len(data)} postings, {n_queries} queries, k={k}  ({time.time() - t0:.0f} s)")


def column(t):
    return doc[indptr[t]:indptr[t + 1]], data[indptr[t]:indptr[t + 1]]


def simulate(query, tiers):
    toks, mult = np.unique(np.asarray(query), return_counts=True)
    toks, mult = toks[df[toks] > 0], mult[df[toks] > 0]
    cols = {int(t): column(int(t)) for t in toks}
    m = {int(t): float(c) for t, c in zip(toks, mult)}
    # pseudo-lists: (ub, token, kind, tau)   kind 0 = whole column, 1 = high tier (v > tau), 2 = low tier (v <= tau)
    pl = []
    for t in cols:
        d_, v_ = cols[t]
        if tiers and len(v_) >= TIER_MIN_DF:
            r = max(16, int(len(v_) * TIER_FRAC))
            tau = float(np.partition(v_, len(v_) - r)[len(v_) - r])
            if (v_ > tau).any():
                pl.append((m[t] * float(v_.max()), t, 1, tau))
                pl.append((m[t] * tau, t, 2, tau))
                continue
        pl.append((m[t] * float(v_.max()), t, 0, 0.0))
    pl.sort(key=lambda x: -x[0])
    n = len(pl)
    # R[i]: bound of a document that is in none of the pseudo-lists before i
    cur, R = {t: 0.0 for t in cols}, [0.0] * (n + 1)
    for i in range(n - 1, -1, -1):
        ub, t, kind, tau = pl[i]
        R[i] = R[i + 1] - cur[t] + ub
        cur[t] = ub
    top = np.zeros(0, np.float32)
    top_docs = np.zeros(0, np.int64)
    theta = 0.0
    streamed = survivors = lookups = 0
    done = {t: [] for t in cols}  # processed tiers of a token: list of (kind, tau)

    def processed(t, val):  # was a posting with this value scored when an earlier pseudo-list was streamed?
        out = np.zeros(len(val), bool)
        for kind, tau in done[t]:
            out |= (val > tau) if kind == 1 else ((val <= tau) if kind == 2 else np.ones(len(val), bool))
        return out

    for i in range(n):
        if R[i] < theta:
            break
        ub, t, kind, tau = pl[i]
        d_, v_ = cols[t]
        if kind == 1:
            sel = v_ > tau
            d_, v_ = d_[sel], v_[sel]
        # own remaining tier must not be counted next to the posting's own value
        own_next = m[t] * tau if kind == 1 else 0.0
        rest = R[i + 1] - own_next
        # other tokens' possible contribution once list i is being streamed
        others = [(w, max([u for (u, tw, kw, _) in pl[i + 1:] if tw == w], default=0.0)) for w in cols if w != t]
        others.sort(key=lambda x: -x[1])
        for c0 in range(0, len(v_), CHUNK):
            dv, vv = d_[c0:c0 + CHUNK], v_[c0:c0 + CHUNK]
            streamed += len(vv)
            keep = (m[t] * vv + rest >= theta)
            if kind == 2:
                keep &= vv <= tau
            dv, vv = dv[keep], vv[keep]
            if len(vv) == 0:
                continue
            survivors += len(vv)
            acc = m[t] * vv.astype(np.float64)
            alive = np.ones(len(vv), bool)
            remaining = sum(c for _, c in others)
            for w, c in others:
                remaining -= c
                if not alive.any():
                    break
                dw, vw = cols[w]
                idx = np.flatnonzero(alive)
                lookups += len(idx)
                pos = np.searchsorted(dw, dv[idx])
                found = (pos < len(dw)) & (dw[np.minimum(pos, len(dw) - 1)] == dv[idx])
                val = np.where(found, vw[np.minimum(pos, len(dw) - 1)], 0.0)
                already = found & processed(w, val)
                acc[idx] += m[w] * val
                alive[idx[already]] = False
                alive &= acc + remaining >= theta
            dv, sc = dv[alive], acc[alive].astype(np.float32)
            if len(sc):
                top = np.concatenate([top, sc])
                top_docs = np.concatenate([top_docs, dv])
                if len(top) > k:
                    o = np.argsort(-top, kind="stable")[:k]
                    top, top_docs = top[o], top_docs[o]
                if len(top) >= k:
                    theta = max(theta, float(top.min()))
        done[t].append((kind, tau))
    return streamed, survivors, lookups, np.sort(top)[::-1]


def brute(query):
    acc = {}
    for t in query:
        d_, v_ = column(int(t))
        for a, b in zip(d_.tolist(), v_.tolist()):
            acc[a] = acc.get(a, 0.0) + b
    return np.sort(np.array(list(acc.values()), np.float32))[::-1][:k]


tot = {False: [0, 0, 0], True: [0, 0, 0]}
postings = 0
checked = 0
t0 = time.time()
for qi in range(n_queries):
    q = qs.tokens[qs.ptr[qi]:qs.ptr[qi + 1]].tolist()
    postings += int(sum(df[t] for t in q))
    res = {}
    for tiers in (False, True):
        s_, v_, l_, top = simulate(q, tiers)
        tot[tiers][0] += s_
        tot[tiers][1] += v_
        tot[tiers][2] += l_
        res[tiers] = top
    if len(res[False]) == k and len(res[True]) == k:
        assert np.allclose(res[False], res[True], rtol=1e-5), (qi, res)
        if qi % 50 == 0 and sum(df[t] for t in q) < 300_000:
            assert np.allclose(brute(q), res[True], rtol=1e-5), qi
            checked += 1
print(f"{n_queries} queries, {postings} postings of the batch  ({time.time() - t0:.0f} s, {checked} queries checked against a brute-force sum)")
for tiers in (False, True):
    s_, v_, l_ = tot[tiers]
    print(f"{'impact tiers' if tiers else 'column maxima'}: streamed {s_} ({100 * s_ / postings:.1f} % of the postings), survivors {v_}, lookups {l_}")
