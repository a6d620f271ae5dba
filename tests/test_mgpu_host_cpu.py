"""CPU test of the multi-GPU host logic with two gloo ranks: every rank derives its column slice, the per-slice
bests are all-gathered and merged with the reference's tie-break, exactly as bench.py / tests/mgpu_check.py do on
NCCL.  (The device side of the chain needs GPUs: tests/mgpu_check.py under torchrun.)"""
import os
import sys

import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    b200 = load_package()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1001
    j0, j1 = b200.column_slice(n, rank, world)
    local_best = (100, 5, j0 + 3) if rank == 0 else (100, 5, j0 + 1)      # same score and row: smaller column wins
    bests = [None] * world
    dist.all_gather_object(bests, local_best)
    spans = [None] * world
    dist.all_gather_object(spans, (j0, j1))
    q.put((rank, b200.merge_best(bests), spans))
    dist.destroy_process_group()


def test_two_rank_gloo_slices_and_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _rank, best, spans in res:
        assert best == (100, 5, 3)
        assert spans == [(0, 500), (500, 1001)]
