"""Shared comparison of an Engine round against the oracle round (bit-exact, every output)."""
import numpy as np


def assert_round_equal(eng_res, fit_rows, score_rows, orc, check_sort=True):
    np.testing.assert_array_equal(eng_res.prefilter, orc.prefilter, err_msg="prefilter")
    np.testing.assert_array_equal(eng_res.new_denied, orc.new_denied, err_msg="new_denied")
    assert eng_res.max_group == orc.max_group, ("max_group", eng_res.max_group, orc.max_group)
    assert eng_res.max_finished == orc.max_finished
    np.testing.assert_array_equal(eng_res.feasible_count, orc.feasible_count, err_msg="feasible_count")
    if fit_rows is not None and orc.fit_bitmap is not None:
        np.testing.assert_array_equal(fit_rows, orc.fit_bitmap, err_msg="fit bitmap")
    if score_rows is not None and orc.score is not None:
        np.testing.assert_array_equal(score_rows, orc.score, err_msg="score matrix")
    np.testing.assert_array_equal(eng_res.best_node, orc.best_node, err_msg="best_node")
    np.testing.assert_array_equal(eng_res.best_score, orc.best_score, err_msg="best_score")
    np.testing.assert_array_equal(eng_res.admit, orc.admit, err_msg="admit")
    np.testing.assert_array_equal(eng_res.admit_bitmap, orc.admit_bitmap, err_msg="admit bitmap")
    if check_sort:
        np.testing.assert_array_equal(eng_res.order, orc.order, err_msg="order")
        np.testing.assert_array_equal(eng_res.rank, orc.rank, err_msg="rank")


def run_and_compare(pkg, oracle, snap, score=True, check_sort=True):
    eng = pkg.Engine(snap.lanes, 0, fit_bitmap=True, score=score)
    try:
        eng.upload(snap)
        res = eng.evaluate()
        fit = eng.fit_rows()
        sc = eng.score_rows() if score else None
    finally:
        eng.close()
    orc = oracle.round(snap, want_bitmap=True, want_score=score)
    assert not orc.ref_panic
    assert_round_equal(res, fit, sc, orc, check_sort)
    return res, orc
