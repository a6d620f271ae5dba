#!/usr/bin/env python3
"""Development tool: digest the per-job timestamps a chained run leaves in $B200_TRACE_DIR (engine.cu, B200_TRACE_DIR).

For every GPU: how many jobs were running over time, how long jobs waited in the queue (pushed -> popped), how long a
job needs to its first publication, and the activity profile of the whole chain in 50 slices of the run."""
import glob, os, sys
import numpy as np

d = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("B200_TRACE_DIR", ".")
files = sorted(glob.glob(os.path.join(d, "trace_rank*.bin")))
ranks = []
for fn in files:
    raw = np.fromfile(fn, dtype=np.uint64)
    S, K, world, rank, C, cmax, dur_ns = (int(x) for x in raw[:7])
    t = raw[8:].reshape(-1, 4).astype(np.int64)
    ranks.append(dict(S=S, K=K, world=world, rank=rank, t=t, dur_ns=dur_ns))
    ok = t[:, 1] > 0
    t0 = t[ok, 1].min(); t1 = t[ok, 3].max()
    wait = (t[ok, 1] - t[ok, 0])[t[ok, 0] > 0] / 1e3
    run = (t[ok, 3] - t[ok, 1]) / 1e3
    first = (t[ok, 2] - t[ok, 1])[t[ok, 2] > 0] / 1e3
    print(f"rank {rank}: {ok.sum()} jobs ({S} strips x {K} chunks), span {(t1 - t0) / 1e6:.1f} ms (kernel {dur_ns / 1e6:.1f} ms); "
          f"queue wait us p50/p90/p99/max = {np.percentile(wait, 50):.1f}/{np.percentile(wait, 90):.1f}/{np.percentile(wait, 99):.1f}/{wait.max():.1f}; "
          f"job run ms p50/p90/max = {np.percentile(run, 50) / 1e3:.2f}/{np.percentile(run, 90) / 1e3:.2f}/{run.max() / 1e3:.2f}; "
          f"first publication us p50/p90 = {np.percentile(first, 50):.1f}/{np.percentile(first, 90):.1f}")
    # running jobs over time (50 slices)
    edges = np.linspace(t0, t1, 51)
    act = []
    for a, b in zip(edges[:-1], edges[1:]):
        ov = np.clip(np.minimum(t[ok, 3], b) - np.maximum(t[ok, 1], a), 0, None).sum() / (b - a)
        act.append(ov)
    print("   running jobs per slice:", " ".join(f"{int(x)}" for x in act))
if ranks:
    # per strip: total time from first pop to last finish, and the sum of queue waits along the strip (same-clock pushes only)
    S = ranks[0]["S"]
    print("strip-level: (rank 0 clock) strip start/end in ms for a few strips")
    t = ranks[0]["t"]; K = ranks[0]["K"]
    base = t[t[:, 1] > 0, 1].min()
    for r in [0, S // 4, S // 2, 3 * S // 4, S - 1]:
        js = [k * S + r for k in range(K)]
        st = [(t[j, 1] - base) / 1e6 for j in js if t[j, 1] > 0]
        en = [(t[j, 3] - base) / 1e6 for j in js if t[j, 3] > 0]
        if st:
            print(f"   strip {r}: first pop {min(st):.1f} ms, last finish {max(en):.1f} ms")
