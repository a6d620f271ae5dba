"""GPU parity of b200_match_last_column (device goal matcher) against the oracle's restatement of
AlignerUtils::matchColumn (C/libmasa/utils/AlignerUtils.cpp:50-107): first k wins, H+H before E+E+open at the same k,
an overshoot before any hit is an error (type -1 / -2)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _case(rng, n, goal, plant):
    buf = np.zeros(n, O.CELL); base = np.zeros(n, O.CELL)
    buf["h"] = rng.integers(-50, 20, n); base["h"] = rng.integers(-50, 20, n)
    buf["x"] = rng.integers(-60, 10, n); base["x"] = rng.integers(-60, 10, n)
    for kind, k in plant:
        if kind == "match":
            base["h"][k] = goal - buf["h"][k]
        elif kind == "gap":
            base["x"][k] = goal - 3 - buf["x"][k]
        elif kind == "over":
            base["h"][k] = goal - buf["h"][k] + 7
        elif kind == "overgap":
            base["x"][k] = goal - 3 - buf["x"][k] + 5
    return buf, base


@pytest.mark.parametrize("n,plant", [
    (1, []), (1, [("match", 0)]), (700, []), (700, [("match", 333)]), (700, [("gap", 12), ("match", 500)]),
    (5000, [("match", 4999)]), (5000, [("gap", 4100), ("match", 4100)]), (5000, [("over", 77), ("match", 300)]),
    (5000, [("match", 300), ("over", 900)]), (3000, [("overgap", 1500)]), (1024, [("gap", 1023)]),
])
def test_match_last_column_matches_oracle(b200, aligner, n, plant):
    rng = np.random.default_rng(n * 31 + len(plant))
    goal = 100                                  # random sums stay below 40: only planted events fire
    buf, base = _case(rng, n, goal, plant)
    want = O.match_column(buf, base, goal)
    got = aligner.match_last_column(buf, base, goal)
    assert got["found"] == want["found"] and got["k"] == want["k"] and got["type"] == want["type"]
    if want["found"]:
        assert got["score"] == want["score"]


def test_match_does_not_disturb_a_chunked_alignment(b200):
    """The matcher has its own scratch: calling it between two B200_CONT_CHUNK launches must not clobber the left
    border of the running partition."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import synth
    a, b = synth.make_pair(6000, 5000, [(500, 5000)], 0.05, 0.01, 0.01, 0, 3)
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    r0 = al.align_partition(recurrence=b200.NEEDLEMAN_WUNSCH, first_row_init=b200.INIT_GAPS, first_col_init=b200.INIT_GAPS,
                            want_last_column=True, want_best_score=False)
    rng = np.random.default_rng(1)
    buf, base = _case(rng, 9000, 100, [("match", 8000)])
    assert al.match_last_column(buf, base, 100)["k"] == 8000
    r1 = al.align_partition(recurrence=b200.NEEDLEMAN_WUNSCH, first_row_init=b200.INIT_GAPS, first_col_init=b200.INIT_GAPS,
                            want_last_column=True, want_best_score=False)
    assert np.array_equal(r0["last_column"], r1["last_column"])
    o = O.full_matrix(a, b, O.NW, first_row_type=O.INIT_GAPS, first_col_type=O.INIT_GAPS)
    assert np.array_equal(r1["last_column"], o["last_col"])
    al.close()
