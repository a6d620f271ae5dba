#!/usr/bin/env python3
"""Development tool (one GPU): throughput of the chain machinery under starvation.

Runs the same stage-1 problem (a) through the single-GPU persistent kernel and (b) through the block-cyclic chain with
the GPU as its own neighbour (world = 1) or as several ranks sharing the device, and prints GCUPS over the computed
cells plus the share of warp time spent computing.  `--strips-per-sm` sizes the rows so that only that many strips
exist per SM (a multi-GPU run at N GPUs leaves each GPU with about 1/N of the front).

  python tools/chain_perf.py --strips-per-sm 8 --cols 3000000 --chunk 65536
  python tools/chain_perf.py --config cfg3 --scale 0.1 --prune --chunk 0
"""
import argparse, importlib.util, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth
spec = importlib.util.spec_from_file_location("masa_cudalign_b200", os.path.join(ROOT, "masa-cudalign_b200", "__init__.py"))
b200 = importlib.util.module_from_spec(spec); spec.loader.exec_module(b200)

ap = argparse.ArgumentParser()
ap.add_argument("--strips-per-sm", type=float, default=8)
ap.add_argument("--cols", type=int, default=3_000_000)
ap.add_argument("--config", default=None)
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--prune", action="store_true")
ap.add_argument("--chunk", type=int, default=65536)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--modes", default="single,chain")
ap.add_argument("--warps-per-sm", type=int, default=0)
ap.add_argument("--group", type=int, default=0, help="mode 'group': this many ranks share device 0, each with 16/N warps per SM")
args = ap.parse_args()
if args.config:
    a, b = synth.make_config(args.config, args.scale)
else:
    m = int(args.strips_per_sm * 148) * 1024
    a, b = synth.make_pair(m, args.cols, [(int(m * 0.1), int(m * 0.9))], 0.05, 0.01, 0.01, 0, 5)
m, n = a.size, b.size
out = {"m": m, "n": n, "prune": args.prune, "chunk": args.chunk}
for mode in args.modes.split(","):
    if mode == "group":
        os.environ["B200_GROUP_WARPS_PER_SM"] = str(16 // args.group)
        g = b200.Group([0] * args.group, m, n, args.chunk)
        g.set_sequences(a, b)
        for rep in range(args.reps):
            best = g.align_partition(use_callbacks=False, prune=args.prune, chunk_cols=args.chunk)
        per = g.rank_results()
        out[mode] = {"best": best["best"], "ms": round(best["device_ms"], 1), "gcups_computed": round(best["cells"] / best["device_ms"] / 1e6, 1),
                     "computed_frac": round(best["cells"] / (m * n), 3), "warp_busy": best["warp_busy"], "chunks": best["chunks"],
                     "per_rank": [(round(p["device_ms"], 1), p["warp_busy"], round(p["cells"] / max(best["cells"], 1), 3)) for p in per]}
        g.close()
        continue
    al = b200.Aligner(warps_per_sm=args.warps_per_sm)
    if mode == "chain":
        al.mgpu_setup(None, 0, 1, m, n, args.chunk)
    al.set_sequences(a, b)
    best = None
    for rep in range(args.reps):
        r = al.align_partition(use_callbacks=False, prune=args.prune, mgpu=(mode == "chain"), chunk_cols=args.chunk)
        best = r
    out[mode] = {"best": best["best"], "ms": round(best["device_ms"], 1), "gcups_computed": round(best["cells"] / best["device_ms"] / 1e6, 1),
                 "computed_frac": round(best["cells"] / (m * n), 3), "warp_busy": best["warp_busy"], "warps": best["warps"], "chunks": best["chunks"]}
    al.close()
print(json.dumps(out))
