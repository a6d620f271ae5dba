"""Run under torchrun on >= 2 GPUs: the column-split chained wavefront (peer-memory border stream) must give the
same best cell and the same last column / last row as one GPU and as the CPU oracle.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402
import oracle_lib as O  # noqa: E402
import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b200 = load_package()
    ok = True
    for (m, n, kern) in [(6000, 5000, b200.KERNEL_S16X2), (6000, 5000, b200.KERNEL_S32), (40000, 30011, b200.KERNEL_S16X2), (300000, 200000, b200.KERNEL_S16X2)]:
        a, b = synth.make_pair(m, n, [(m // 10, m * 8 // 10)], 0.05, 0.01, 0.01, 0, 77)
        al = b200.Aligner(device=local, kernel=kern)
        al.mgpu_setup(dist, rank, world, m)
        al.set_sequences(a, b)
        j0, j1 = n * rank // world, n * (rank + 1) // world
        for rep in range(2):                      # twice: the exchange block must re-arm correctly
            torch.cuda.synchronize(); dist.barrier()
            r = al.align_partition(0, j0, m, j1, want_best_score=True, want_last_column=(rank == world - 1), want_last_row=True, mgpu=True)
            torch.cuda.synchronize(); dist.barrier()
            bests = [None] * world
            dist.all_gather_object(bests, tuple(r["best"]))
            best = max(bests, key=lambda s: (s[0], -s[1], -s[2]))
            if rank == world - 1:
                if m * n <= 2_000_000_000:
                    o = O.full_matrix(a, b, O.SW, row_ids=[m - 1])
                    good = best == o["best"] and np.array_equal(r["last_column"], o["last_col"]) and np.array_equal(r["rows"][m][1:], o["rows"][m - 1][1 + j0:])
                else:
                    al1 = b200.Aligner(device=local, kernel=kern)
                    al1.set_sequences(a, b)
                    r1 = al1.align_partition(want_best_score=True, want_last_column=True, want_last_row=True)
                    good = best == r1["best"] and np.array_equal(r["last_column"], r1["last_column"]) and np.array_equal(r["rows"][m][1:], r1["rows"][m][1 + j0:])
                    al1.close()
                print(f"mgpu {m}x{n} kernel={kern} world={world} rep={rep}: best={best} {'OK' if good else 'MISMATCH'}", flush=True)
                ok = ok and good
        al.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
