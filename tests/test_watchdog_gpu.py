"""The spin-wait watchdog of the strip chain is a TIME without progress, sized from the chunk width and the number of
GPUs (csrc/strip_common.cuh Watchdog, engine.cu watchdog_ns) -- not a spin count: a neighbour that is legitimately slow
for many seconds (a wide slice on another GPU) must not abort the run, a dependency that never advances must."""
import os
import sys

import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chained", [False, True])
def test_ten_second_upstream_delay_does_not_trip(b200, monkeypatch, chained):
    monkeypatch.setenv("B200_TEST_DELAY_MS", "10000")      # the first strip sleeps 10 s: every other resident warp waits on it
    monkeypatch.delenv("B200_WATCHDOG_S", raising=False)
    a, b = synth.make_pair(40000, 9000, [(2000, 30000)], 0.05, 0.01, 0.01, 0, 5)
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    if chained:
        al.mgpu_setup(None, 0, 1, a.size, b.size, 1000)
    al.set_sequences(a, b)
    r = al.align_partition(use_callbacks=False, mgpu=chained, chunk_cols=1000 if chained else 0)
    assert r["device_ms"] > 9000
    assert r["best"] == O.full_matrix(a, b, O.SW, want_last_col=False)["best"]
    al.close()


def test_stuck_dependency_is_reported(b200, monkeypatch):
    monkeypatch.setenv("B200_TEST_DELAY_MS", "6000")
    monkeypatch.setenv("B200_WATCHDOG_S", "1.5")           # shorter than the delay: the waiters must give up with an error
    a, b = synth.make_pair(40000, 9000, [(2000, 30000)], 0.05, 0.01, 0.01, 0, 5)
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    with pytest.raises(b200.B200Error, match="watchdog"):
        al.align_partition(use_callbacks=False)
    monkeypatch.delenv("B200_TEST_DELAY_MS")
    monkeypatch.delenv("B200_WATCHDOG_S")
    r = al.align_partition(use_callbacks=False)            # the handle stays usable
    assert r["best"] == O.full_matrix(a, b, O.SW, want_last_col=False)["best"]
    al.close()
