#!/usr/bin/env python3
"""Quick device-timed GCUPS probe of b200_align_partition (development aid, not the benchmark)."""
import argparse, importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth
spec = importlib.util.spec_from_file_location("masa_cudalign_b200", os.path.join(ROOT, "masa-cudalign_b200", "__init__.py"))
b200 = importlib.util.module_from_spec(spec); spec.loader.exec_module(b200)

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="100000,300000,1000000")
ap.add_argument("--kernel", default="auto")
ap.add_argument("--wps", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--nw", action="store_true")
ap.add_argument("--nobest", action="store_true")
ap.add_argument("--prune", action="store_true")
ap.add_argument("--opts", default="", help="comma list of B200_OPT values to run each size with")
args = ap.parse_args()
k = {"auto": b200.KERNEL_AUTO, "s32": b200.KERNEL_S32, "s16x2": b200.KERNEL_S16X2}[args.kernel]
al = b200.Aligner(kernel=k, warps_per_sm=args.wps)
for sz in args.sizes.split(","):
    if "x" in sz:
        m, n = (int(v) for v in sz.split("x"))
    else:
        m = n = int(sz)
    a, b = synth.make_pair(m, n, [(m // 10, m * 9 // 10)], 0.05, 0.01, 0.01, 0, 1234)
    al.set_sequences(a, b)
    for rep_i in range(args.reps * max(1, len(args.opts.split(",")) if args.opts else 1)):
        rep = rep_i % args.reps
        if args.opts:
            os.environ["B200_OPT"] = args.opts.split(",")[rep_i // args.reps]
        t0 = time.time()
        r = al.align_partition(recurrence=b200.NEEDLEMAN_WUNSCH if args.nw else b200.SMITH_WATERMAN,
                               first_row_init=b200.INIT_GAPS if args.nw else 0, first_col_init=b200.INIT_GAPS if args.nw else 0,
                               want_best_score=not (args.nw or args.nobest), prune=args.prune, use_callbacks=False)
        wall = time.time() - t0
        print(f"opt={os.environ.get('B200_OPT', '-'):>3} {m}x{n} kernel={r['kernel_used']} strips={r['strips']} best={r['best']} dev_ms={r['device_ms']:.2f} "
              f"GCUPS={m*n/r['device_ms']/1e6:.1f} computed={r['cells']/(m*n)*100:.1f}% wall_ms={wall*1e3:.1f}", flush=True)
