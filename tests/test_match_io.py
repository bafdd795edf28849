"""Match file formats: the reference's dense `.npz` layout (match_line_pairs.py:94-104)
and the compact index layout, with converters; plus the oracle's restatement of the training-side
matcher (evaluations/matcher.py) against the committed outputs of the reference functions."""
import os

import numpy as np
import pytest

from linetr_b200 import match_io as mio
from oracle import linetr_oracle as orc
from tests import helpers as H


def _pair():
    rng = np.random.Generator(np.random.PCG64(4))
    kp0, kp1 = rng.uniform(0, 640, (30, 2)).astype(np.float32), rng.uniform(0, 640, (25, 2)).astype(np.float32)
    kl0, kl1 = rng.uniform(0, 480, (12, 2, 2)).astype(np.float32), rng.uniform(0, 480, (9, 2, 2)).astype(np.float32)
    dist_p = rng.uniform(0, 2, (30, 25)).astype(np.float32)
    dist_l = rng.uniform(0, 2, (12, 9)).astype(np.float32)
    mat_p = orc.nn_matcher_distmat(dist_p[None], 0.7, True)[0]
    mat_l = orc.nn_matcher_distmat(dist_l[None], 0.8, True)[0]
    return kp0, kp1, kl0, kl1, mat_p, mat_l, dist_p, dist_l


def test_reference_layout_has_exactly_the_reference_keys(tmp_path):
    kp0, kp1, kl0, kl1, mat_p, mat_l, dist_p, dist_l = _pair()
    f = tmp_path / "a_b_matches.npz"
    mio.save_matches_npz(f, kp0, kp1, kl0, kl1, mio.dense_to_indices(mat_p), mio.dense_to_indices(mat_l), dist_p, dist_l)
    z = np.load(f)
    assert sorted(z.files) == sorted(mio.REFERENCE_KEYS)      # match_line_pairs.py:100-103
    assert z["matches_l"].dtype == np.float64 and np.array_equal(z["matches_l"], mat_l)
    assert np.array_equal(z["matches_p"], mat_p) and np.array_equal(z["match_confidence_l"], dist_l)
    # what the reference does with the file afterwards (match_line_pairs.py:107-113)
    m = np.where(z["matches_l"] > 0)
    assert np.array_equal(m[1], mio.dense_to_indices(mat_l)[m[0]])


def test_compact_round_trip_and_size(tmp_path):
    kp0, kp1, kl0, kl1, mat_p, mat_l, dist_p, dist_l = _pair()
    ref, cmp_ = tmp_path / "ref.npz", tmp_path / "compact.npz"
    mio.save_matches_npz(ref, kp0, kp1, kl0, kl1, mat_p, mat_l, dist_p, dist_l)
    mio.save_matches_npz(cmp_, kp0, kp1, kl0, kl1, mat_p, mat_l, dist_p, dist_l, compact=True)
    a, b = mio.load_matches_npz(ref), mio.load_matches_npz(cmp_)
    for k in ("keypoints0", "keypoints1", "keylines0", "keylines1", "matches_p", "matches_l", "matches_p_idx", "matches_l_idx"):
        assert np.array_equal(a[k], b[k]), k
    rows = np.nonzero(a["matches_l_idx"] >= 0)[0]
    assert np.array_equal(b["match_confidence_l"][rows, a["matches_l_idx"][rows]], dist_l[rows, a["matches_l_idx"][rows]])
    assert np.isnan(b["match_confidence_l"]).sum() == dist_l.size - len(rows)
    assert os.path.getsize(cmp_) < os.path.getsize(ref)
    with pytest.raises(ValueError):
        mio.save_matches_npz(ref, kp0, kp1, kl0, kl1, mat_p, mat_l)          # reference layout needs the dense distances


def test_index_dense_converters_edge_cases():
    assert mio.dense_to_indices(np.zeros((1, 3, 0))).tolist() == [-1, -1, -1]
    assert mio.indices_to_dense(np.array([-1, 2], np.int32), 3).tolist() == [[0, 0, 0], [0, 0, 1]]
    assert mio.indices_to_dense(np.zeros(0, np.int32), 4).shape == (0, 4)


def test_oracle_eval_matcher_vs_reference_outputs():
    npz = dict(np.load(os.path.join(H.GOLDEN_DIR, "eval_matcher_outputs.npz")))
    d0, d1 = npz["eval_desc0"], npz["eval_desc1"]
    for mutual in (False, True):
        got = orc.eval_nn_matcher_batches(d0, d1, 0.9, mutual)
        assert got.shape == (3, 151, 132) and np.array_equal(got, npz[f"eval_batches_m{int(mutual)}"])
        assert np.array_equal(got[0, :-1, :-1], npz[f"eval_single_m{int(mutual)}"])
        assert np.array_equal(orc.nn_matcher_distmat(npz["eval_score_in"][None], 0.5, mutual)[0], npz[f"eval_score_m{int(mutual)}"])
