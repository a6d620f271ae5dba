#!/usr/bin/env python3
"""Stage-1 SW (best score + end coordinate, block pruning) on a named BASELINE config, single GPU. Development/record tool."""
import importlib.util, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth
spec = importlib.util.spec_from_file_location("masa_cudalign_b200", os.path.join(ROOT, "masa-cudalign_b200", "__init__.py"))
b200 = importlib.util.module_from_spec(spec); spec.loader.exec_module(b200)
name = sys.argv[1]; scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
t0 = time.time(); a, b = synth.make_config(name, scale); tg = time.time() - t0
al = b200.Aligner()
t0 = time.time(); al.set_sequences(a, b); tu = time.time() - t0
out = {"config": name, "scale": scale, "m": int(a.size), "n": int(b.size), "gen_s": tg, "upload_s": tu}
for prune in (True,):
    t0 = time.time()
    r = al.align_partition(prune=prune, use_callbacks=False)
    out["prune" if prune else "noprune"] = {"best": r["best"], "device_ms": r["device_ms"], "wall_s": time.time() - t0,
        "gcups": a.size * b.size / r["device_ms"] / 1e6, "computed_frac": r["cells"] / (a.size * b.size), "strips": r["strips"]}
print(json.dumps(out))
