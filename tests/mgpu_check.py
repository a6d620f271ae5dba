"""Run under torchrun on >= 2 GPUs (tests/test_mgpu_gpu.py does): the block-cyclic chained wavefront over NVLink peer
memory must give, on every case, the same best cell, special rows, last row and last column as the CPU oracle (up to
2e9 cells) or as a single-GPU run of the same library (above), with block pruning on and off, SW and NW, both kernels,
automatic / narrow / one-slice-per-GPU chunking, and on repeated calls with different sequences (re-arm of the exchange
blocks and of the shared running best).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

MGPU_CHECK_BACKEND=gloo: the same script where no GPU exists, one process per EMULATED device (tests/emu/, B200_LIB pointing at
the emulation build, B200_EMU_SHM=1 so that the exchange blocks are shared between the processes): tests/test_emu_cpu.py.
"""
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

SW, NW = 1, 0

# (name, m, n, kernel, recurrence, prune, chunk_cols, artefacts, homology lo/hi of the repetitions)
CASES = [
    ("sw_small_s16", 6000, 5000, "s16x2", SW, False, 700, True, [(0.1, 0.8)]),
    ("sw_small_s32", 6000, 5000, "s32", SW, False, 700, True, [(0.1, 0.8)]),
    ("sw_slices", 30000, 26011, "s16x2", SW, False, -1, True, [(0.1, 0.9)]),
    ("sw_auto_chunks", 40000, 300000, "s16x2", SW, False, 0, True, [(0.1, 0.9)]),
    ("sw_pruned_narrow", 45000, 44000, "s16x2", SW, True, 1500, False, [(0.05, 0.95), (0.45, 0.55), (0.2, 0.7)]),
    ("sw_pruned_slices", 45000, 44000, "s16x2", SW, True, -1, False, [(0.05, 0.95), (0.4, 0.6)]),
    ("nw_global", 20000, 21000, "s16x2", NW, False, 3000, True, [(0.0, 1.0)]),
    ("nw_global_s32", 9000, 8000, "s32", NW, False, 1000, True, [(0.0, 1.0)]),
    ("sw_big_pruned", 600000, 500000, "s16x2", SW, True, 0, False, [(0.02, 0.98), (0.3, 0.5)]),
    ("sw_big_unpruned", 300000, 400000, "s16x2", SW, False, 16384, False, [(0.1, 0.9)]),
]


def assemble(pieces_per_rank, chunks, first_cell_rank=0):
    """Rows arrive per rank as [first-column cell (rank 0)] + its chunks in column order: rebuild the whole row."""
    pos = [1 if r == first_cell_rank else 0 for r in range(len(pieces_per_rank))]
    out = [pieces_per_rank[first_cell_rank][:1]]
    for j0, j1, owner in chunks:
        out.append(pieces_per_rank[owner][pos[owner]:pos[owner] + (j1 - j0)])
        pos[owner] += j1 - j0
    return np.concatenate(out)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    gloo = os.environ.get("MGPU_CHECK_BACKEND", "nccl") == "gloo"
    if gloo:
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def sync():
        if not gloo:
            torch.cuda.synchronize()
        dist.barrier()

    b200 = load_package()
    only = os.environ.get("MGPU_CASES")
    ok = True
    for name, m, n, kern, rec, prune, chunk, arte, reps in CASES:
        if only and name not in only.split(","):
            continue
        kernel = b200.KERNEL_S32 if kern == "s32" else b200.KERNEL_S16X2
        al = b200.Aligner(device=local, kernel=kernel)
        al.mgpu_setup(dist, rank, world, m, n, chunk)
        chunks = b200.chain_chunks(n, world, chunk)
        last_owner = chunks[-1][2]
        for rep, (lo, hi) in enumerate(reps):
            a, b = synth.make_pair(m, n, [(int(m * lo), int(m * hi))], 0.05, 0.01, 0.01, 0, 77 + rep)
            al.set_sequences(a, b)
            sync()
            kw = dict(recurrence=rec, want_best_score=(rec == SW), prune=prune, mgpu=True, chunk_cols=chunk)
            if rec == NW:
                kw.update(first_row_init=b200.INIT_GAPS, first_col_init=b200.INIT_GAPS)
            if arte:
                kw.update(want_last_row=True, want_last_column=True, want_special_rows=True, special_row_interval=1000)
            else:
                kw.update(use_callbacks=False)
            r = al.align_partition(0, 0, m, n, **kw)
            sync()
            got = [None] * world
            dist.all_gather_object(got, dict(best=tuple(r["best"]), cells=r["cells"], rows=r["rows"] if arte else {},
                                             last_column=r["last_column"] if arte else None))
            if rank != 0:
                continue
            best = b200.merge_best([g["best"] for g in got]) if rec == SW else None
            cells = sum(g["cells"] for g in got)
            good = True
            if m * n <= 2_000_000_000:
                ids = sorted(i for i in got[0]["rows"] if i != m) if arte else []
                o = O.full_matrix(a, b, rec, first_row_type=O.INIT_GAPS if rec == NW else O.INIT_ZEROES,
                                  first_col_type=O.INIT_GAPS if rec == NW else O.INIT_ZEROES,
                                  row_ids=[i - 1 for i in ids] + [m - 1], want_last_col=arte)
                if rec == SW:
                    good &= best == o["best"]
                if arte:
                    good &= ids == list(range(8192, m, 8192))
                    for i in ids + [m]:
                        row = assemble([g["rows"].get(i, np.zeros(0, O.CELL)) for g in got], chunks)     # a rank may own no chunk
                        good &= np.array_equal(row, o["rows"][i - 1])
                    good &= np.array_equal(got[last_owner]["last_column"], o["last_col"])
                ref = "oracle"
            else:
                al1 = b200.Aligner(device=local, kernel=kernel)
                al1.set_sequences(a, b)
                r1 = al1.align_partition(want_best_score=True, use_callbacks=False, prune=False)
                al1.close()
                good &= best == tuple(r1["best"])
                ref = "1 GPU, unpruned"
            if prune:
                good &= cells < 0.9 * m * n                     # pruning engaged, or the case proves nothing
            else:
                good &= cells == m * n
            per = [g["cells"] for g in got]
            print(f"mgpu {name} {m}x{n} world={world} rep={rep} chunks={len(chunks)}: best={best} cells={cells / (m * n):.3f} "
                  f"per-GPU max/mean={max(per) * world / max(cells, 1):.2f} vs {ref}: {'OK' if good else 'MISMATCH'}", flush=True)
            ok = ok and bool(good)
        al.close()
    flag = torch.tensor([0 if ok else 1], device="cpu" if gloo else "cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
