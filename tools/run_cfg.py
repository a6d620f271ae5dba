#!/usr/bin/env python3
"""Stage-1 SW (best score + end coordinate, block pruning) on a named BASELINE config, single GPU. Development/record tool."""
import importlib.util, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth
spec = importlib.util.spec_from_file_location("masa_cudalign_b200", os.path.join(ROOT, "masa-cudalign_b200", "__init__.py"))
b200 = importlib.util.module_from_spec(spec); spec.loader.exec_module(b200)
name = sys.argv[1]; scale = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 1.0
modes = [a[2:] for a in sys.argv[2:] if a.startswith("--")] or ["prune"]      # --prune --noprune --s32 (exact int32 kernel, no pruning)
t0 = time.time(); a, b = synth.make_config(name, scale); tg = time.time() - t0
al = b200.Aligner()
t0 = time.time(); al.set_sequences(a, b); tu = time.time() - t0
out = {"config": name, "scale": scale, "m": int(a.size), "n": int(b.size), "gen_s": tg, "upload_s": tu}
for mode in modes:
    t0 = time.time()
    if mode == "s32":
        al.close(); al = b200.Aligner(kernel=b200.KERNEL_S32); al.set_sequences(a, b)
    r = al.align_partition(prune=(mode == "prune"), use_callbacks=False)
    out[mode] = {"best": r["best"], "device_ms": r["device_ms"], "wall_s": time.time() - t0, "kernel": r["kernel_used"],
        "gcups": a.size * b.size / r["device_ms"] / 1e6, "computed_frac": r["cells"] / (a.size * b.size), "strips": r["strips"]}
    print(json.dumps({mode: out[mode]}), file=sys.stderr, flush=True)
print(json.dumps(out))
