"""Tie-aware parity checker. Ties are arbitrary in the reference (argpartition), so ids are
compared as sets within runs of equal score; scores are compared bit-exactly unless a tolerance is given."""
import numpy as np


def load_golden(name):
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    return np.load(os.path.join(here, "golden", name + ".npz"))


def lists_from_flat(tok, ptr):
    return [tok[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]


def check_topk_rows(ids, scores, ref_ids, ref_scores, dense_fn=None, rtol=0.0, what=""):
    """ids/scores: ours (Q,k); ref_*: reference (Q,k), both sorted by score descending.
    dense_fn(q) -> float32[N] oracle dense scores of query q (optional, enables checks iii and iv)."""
    ids, scores, ref_ids, ref_scores = map(np.asarray, (ids, scores, ref_ids, ref_scores))
    assert ids.shape == ref_ids.shape and scores.shape == ref_scores.shape, (what, ids.shape, ref_ids.shape)
    Q, k = ids.shape
    for q in range(Q):
        s, r = scores[q], ref_scores[q]
        # (i) the sorted score vectors match
        if rtol == 0.0:
            assert np.array_equal(s, r), f"{what} q={q}: scores differ\n{s}\n{r}"
        else:
            np.testing.assert_allclose(s, r, rtol=rtol, atol=0, err_msg=f"{what} q={q}")
        if k == 0:
            continue
        assert np.all(s[:-1] >= s[1:]), f"{what} q={q}: row not sorted descending"
        assert len(set(ids[q].tolist())) == k, f"{what} q={q}: duplicate ids {ids[q]}"
        kth = r[-1]
        # (ii) ids strictly above the k-th score are the same set
        assert set(ids[q][s > kth].tolist()) == set(ref_ids[q][r > kth].tolist()), f"{what} q={q}: id sets differ above the k-th score"
        if dense_fn is not None:
            dense = dense_fn(q)
            # (iv) every returned id carries its oracle dense score
            got = dense[ids[q]]
            if rtol == 0.0:
                assert np.array_equal(got, s), f"{what} q={q}: returned scores are not the dense scores of the returned ids"
            else:
                np.testing.assert_allclose(got, s, rtol=rtol, atol=0)
            # (iii) ids at the k-th score really have that score (subset of the oracle's tie set) - implied by (iv)
            # and nothing outside the result beats the k-th score
            mask = np.ones(len(dense), bool)
            mask[ids[q]] = False
            if mask.any():
                assert dense[mask].max() <= s[-1], f"{what} q={q}: a document outside the result beats the k-th score"
