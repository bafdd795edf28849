"""Match files: the reference's `.npz` layout and a compact index format with converters.

`match_line_pairs.py:94-104` of the reference stores, per image pair, DENSE matrices:
`matches_p` / `matches_l` (float64 0/1, [N0,N1] / [K0,K1]) and `match_confidence_p` /
`match_confidence_l` (float32 distance matrices of the same shapes) next to `keypoints0/1` and
`keylines0/1` - 12 bytes per candidate pair of which at most min(N0,N1) carry information.  The
matcher here produces `int32 matches0[K0]` (index in image 1 or -1) and the distance to that
neighbour; `save_matches_npz(..., compact=True)` keeps exactly that, `load_matches_npz` returns the
reference's keys from either layout, so downstream code written against the reference files
(`np.where(matches_l > 0)`, match_line_pairs.py:107) keeps working.
"""
from __future__ import annotations

import numpy as np

REFERENCE_KEYS = ("keypoints0", "keypoints1", "matches_p", "match_confidence_p",
                  "keylines0", "keylines1", "matches_l", "match_confidence_l")
COMPACT_KEYS = ("keypoints0", "keypoints1", "matches_p_idx", "match_score_p",
                "keylines0", "keylines1", "matches_l_idx", "match_score_l", "format")
COMPACT_FORMAT = "linetr_b200.compact.v1"


def dense_to_indices(mat) -> np.ndarray:
    """0/1 match matrix [n0,n1] (at most one 1 per row, as nn_matcher_distmat builds it) -> int32 [n0]."""
    mat = np.asarray(mat)
    if mat.ndim == 3:
        mat = mat[0]
    if mat.shape[1] == 0:
        return np.full(mat.shape[0], -1, dtype=np.int32)
    return np.where(mat.sum(axis=1) > 0, mat.argmax(axis=1), -1).astype(np.int32)


def indices_to_dense(idx, n1: int) -> np.ndarray:
    """int32 [n0] -> float64 0/1 [n0,n1] (the dtype of models/nn_matcher.py:8,29)."""
    idx = np.asarray(idx)
    mat = np.zeros((len(idx), int(n1)))
    rows = np.nonzero(idx >= 0)[0]
    mat[rows, idx[rows]] = 1
    return mat


def row_scores(conf, idx) -> np.ndarray:
    """Distance of every line/point of image 0 to its match (NaN where unmatched) from a dense matrix."""
    conf = np.asarray(conf)
    if conf.ndim == 3:
        conf = conf[0]
    idx = np.asarray(idx)
    out = np.full(len(idx), np.nan, dtype=np.float32)
    rows = np.nonzero(idx >= 0)[0]
    out[rows] = conf[rows, idx[rows]]
    return out


def save_matches_npz(path, keypoints0, keypoints1, keylines0, keylines1, matches_p, matches_l,
                     confidence_p=None, confidence_l=None, compact=False):
    """matches_*: int32 indices [n0] or dense 0/1 matrices.  confidence_*: dense distance matrices
    ([n0,n1]; required for the reference layout) or per-row scores [n0] (enough for the compact one)."""
    as_idx = lambda m: np.asarray(m, dtype=np.int32) if np.asarray(m).ndim == 1 else dense_to_indices(m)
    ip, il = as_idx(matches_p), as_idx(matches_l)
    n1p, n1l = len(keypoints1), len(keylines1)
    if compact:
        def score(c, idx):
            if c is None:
                return np.full(len(idx), np.nan, dtype=np.float32)
            c = np.asarray(c, dtype=np.float32)
            return c if c.ndim == 1 else row_scores(c, idx)
        np.savez(str(path), keypoints0=keypoints0, keypoints1=keypoints1, matches_p_idx=ip, match_score_p=score(confidence_p, ip),
                 keylines0=keylines0, keylines1=keylines1, matches_l_idx=il, match_score_l=score(confidence_l, il),
                 format=np.array(COMPACT_FORMAT))
        return
    if confidence_p is None or confidence_l is None or np.asarray(confidence_p).ndim < 2 or np.asarray(confidence_l).ndim < 2:
        raise ValueError("the reference layout stores the dense distance matrices: pass confidence_p / confidence_l [n0,n1]")
    sq = lambda c: np.asarray(c)[0] if np.asarray(c).ndim == 3 else np.asarray(c)
    np.savez(str(path), keypoints0=keypoints0, keypoints1=keypoints1, matches_p=indices_to_dense(ip, n1p),
             match_confidence_p=sq(confidence_p), keylines0=keylines0, keylines1=keylines1,
             matches_l=indices_to_dense(il, n1l), match_confidence_l=sq(confidence_l))


def load_matches_npz(path) -> dict:
    """-> dict with the reference's keys (REFERENCE_KEYS) plus `matches_p_idx` / `matches_l_idx`.  From a
    compact file the confidence matrices are rebuilt sparsely: the stored distance at every matched entry,
    NaN elsewhere (the reference itself never reads them back)."""
    with np.load(str(path), allow_pickle=False) as z:
        d = {k: z[k] for k in z.files}
    if "format" in d:
        if str(d["format"]) != COMPACT_FORMAT:
            raise ValueError(f"unknown match file format {d['format']!r}")
        out = {k: d[k] for k in ("keypoints0", "keypoints1", "keylines0", "keylines1")}
        for tag, n1 in (("p", len(d["keypoints1"])), ("l", len(d["keylines1"]))):
            idx = d[f"matches_{tag}_idx"]
            out[f"matches_{tag}_idx"] = idx
            out[f"matches_{tag}"] = indices_to_dense(idx, n1)
            conf = np.full((len(idx), n1), np.nan, dtype=np.float32)
            rows = np.nonzero(idx >= 0)[0]
            conf[rows, idx[rows]] = d[f"match_score_{tag}"][rows]
            out[f"match_confidence_{tag}"] = conf
        return out
    missing = [k for k in REFERENCE_KEYS if k not in d]
    if missing:
        raise ValueError(f"not a LineTR match file: missing {missing}")
    d["matches_p_idx"] = dense_to_indices(d["matches_p"])
    d["matches_l_idx"] = dense_to_indices(d["matches_l"])
    return d
