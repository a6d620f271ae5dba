"""CPU tests (no GPU) of the multi-GPU chain's host-side planning and of its scheduling rule.

 * b200_chain_plan / chain_chunks (the real C ABI, host-only): the column chunks tile [0, n), are dealt round-robin, the
   automatic width stays inside its documented window, the capacities cover the jobs, and chunk_cols < 0 reproduces the
   reference's --split arithmetic (C/libmasa/libmasa.cpp:632-635).
 * A model of the on-device dataflow rule (csrc/strip_common.cuh: chain_arm_kernel's start values, chain_notify_below_,
   chain_notify_right_, chain_pop): one 64-bit event word {left events : 32 | top events : 32} per strip and GPU; whoever
   completes the pair pushes the job.  Under random interleavings every job of every GPU is pushed exactly once, pops
   only see jobs whose two inputs exist, and the run always drains -- the property the persistent kernels rely on (a
   resident warp never waits for a job that cannot start).  The device code itself runs in tests/test_chain_gpu.py and
   tests/mgpu_check.py."""
import random

import pytest


@pytest.mark.parametrize("n,world,chunk_cols", [(10_000_000, 8, 0), (25_000_000, 2, 0), (228_000_001, 8, 0), (5000, 4, 0), (100_000, 3, 7000),
                                                (1_000_000, 8, -1), (5, 8, -1)])
def test_chunks_tile_the_columns(b200, n, world, chunk_cols):
    chunks = b200.chain_chunks(n, world, chunk_cols)
    assert chunks[0][0] == 0 and chunks[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:])) and all(j1 > j0 for j0, j1, _ in chunks)
    assert [o for _, _, o in chunks] == [c % world for c in range(len(chunks))]
    plan = b200.chain_plan(12_345, n, world, chunk_cols)
    if chunk_cols < 0:
        want = sorted(set(n * r // world for r in range(world + 1)))          # libmasa.cpp:632-635, empty slices dropped
        assert [c[0] for c in chunks] + [n] == want
    else:
        assert plan["chunks"] == len(chunks) and plan["chunk_cols"] == chunks[0][1] - chunks[0][0]
        widths = {j1 - j0 for j0, j1, _ in chunks[:-1]}
        assert len(widths) <= 1                                                # only the last chunk may be narrower
        if chunk_cols == 0 and len(chunks) > 1:
            w = plan["chunk_cols"]
            assert 32768 <= w <= (1 << 20) and w % 1024 == 0
    assert plan["chunks_per_gpu"] == -(-plan["chunks"] // world)
    assert plan["max_jobs"] == plan["chunks_per_gpu"] * plan["max_strips"] and plan["max_strips"] >= 12_345 // 1024 + 1


def simulate(S, C, world, rng):
    """Event-word model.  Job (r, c): strip r, chunk c, owner c % world, local index k = c // world.  A running job first
    publishes its first columns (top event for (r+1, c)), later finishes (left event for (r, c+1) on the next owner)."""
    left = [[1 if g == 0 else 0 for _ in range(S)] for g in range(world)]      # rank 0: chunk 0 has no left neighbour
    top = [[(1 << 30) if r == 0 else 0 for r in range(S)] for _ in range(world)]   # strip 0 has no strip above
    queue = [[] for _ in range(world)]
    queue[0].append((0, 0))                                                    # pre-pushed by chain_arm_kernel
    pushed = {(0, 0)}
    published, finished = set(), set()
    running = []                                                               # [job, phase] phase 0 = popped, 1 = published
    total = S * C

    def push(g, r, k):
        c = k * world + g
        assert (r, c) not in pushed, f"job {(r, c)} pushed twice"
        pushed.add((r, c))
        queue[g].append((r, c))

    while len(finished) < total:
        moves = [("pop", g) for g in range(world) if queue[g]] + [("step", i) for i in range(len(running))]
        assert moves, f"deadlock: {len(finished)} of {total} jobs finished"
        kind, x = rng.choice(moves)
        if kind == "pop":
            r, c = queue[x].pop(0)
            assert c == 0 or (r, c - 1) in finished, "popped a job without its left border"
            assert r == 0 or (r - 1, c) in published, "popped a job whose top border has not started"
            running.append([(r, c), 0])
            continue
        (r, c), phase = running[x]
        g, k = c % world, c // world
        if phase == 0:                                                         # chain_notify_below_
            published.add((r, c))
            running[x][1] = 1
            if r + 1 < S:
                old_top, old_left = top[g][r + 1], left[g][r + 1]
                top[g][r + 1] += 1
                if old_top == k and old_left >= k + 1:
                    push(g, r + 1, k)
        else:                                                                  # chain_notify_right_
            finished.add((r, c))
            running.pop(x)
            if c + 1 < C:
                gn, kn = (c + 1) % world, (c + 1) // world
                old_top, old_left = top[gn][r], left[gn][r]
                left[gn][r] += 1
                if old_left == kn and old_top >= kn + 1:
                    push(gn, r, kn)
    assert len(pushed) == total and not any(queue)


@pytest.mark.parametrize("S,C,world", [(1, 1, 1), (5, 1, 1), (1, 7, 2), (6, 9, 1), (7, 8, 2), (9, 13, 4), (12, 16, 8), (3, 5, 8), (20, 3, 2)])
def test_every_job_is_scheduled_exactly_once(S, C, world):
    for seed in range(20):
        simulate(S, C, world, random.Random(1000 * S + 10 * C + world + seed))
