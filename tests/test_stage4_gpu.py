"""GPU parity of the batched stage-4 split (b200_stage4 / b200_stage4_round) against the reference's crosspoint_04
files (tests/golden, produced by the reference's own stage 4) and against the C restatement round by round."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_runs.json")))


def _pair(g):
    return synth.make_pair(g["m"], g["n"], [tuple(g["homology"])], g["p_s"], g["p_d"], g["p_i"], 0, g["seed"])


@pytest.mark.parametrize("kernel", ["s16x2", "s32"])
@pytest.mark.parametrize("name", sorted(GOLD))
def test_stage4_matches_reference_files(b200, name, kernel):
    g = GOLD[name]
    a, b = _pair(g["generator"])
    al = b200.Aligner(kernel=b200.KERNEL_S16X2 if kernel == "s16x2" else b200.KERNEL_S32)
    al.set_sequences(a, b)
    src = sorted(k for k in g["crosspoints"] if k.startswith("crosspoint_03"))[-1]
    got = al.stage4(O.golden_points(g["crosspoints"][src]), 16)
    exp = O.golden_points(g["crosspoints"]["crosspoint_04.00"])
    assert got.size == exp.size
    assert np.array_equal(got, exp)
    al.close()


@pytest.mark.parametrize("max_part", [16, 100, 1])
def test_stage4_rounds_match_oracle(b200, max_part):
    g = GOLD["sw_12k_rows"]
    a, b = _pair(g["generator"])
    al = b200.Aligner()
    al.set_sequences(a, b)
    src = sorted(k for k in g["crosspoints"] if k.startswith("crosspoint_03"))[-1]
    pts = O.golden_points(g["crosspoints"][src])                        # stage-3 output: partitions of a few thousand rows
    for _round in range(12):
        if O.largest_partition(pts) <= max_part:
            break
        mids = al.stage4_round(pts, max_part)
        nxt, changed = O.stage4_round(a, b, pts, max_part)
        merged = [pts[0]]
        for k in range(1, pts.size):
            if mids[k]["type"] != -1 and (mids[k]["i"] != pts[k - 1]["i"] or mids[k]["j"] != pts[k - 1]["j"]):
                merged.append(mids[k])
            merged.append(pts[k])
        merged = np.array(merged, dtype=O.XPOINT)
        assert np.array_equal(merged, nxt), f"round {_round}"
        pts = nxt
        if not changed:
            break
    al.close()
