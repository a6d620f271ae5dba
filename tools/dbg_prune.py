import importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, synth
spec = importlib.util.spec_from_file_location("masa_cudalign_b200", os.path.join(ROOT, "masa-cudalign_b200", "__init__.py"))
b200 = importlib.util.module_from_spec(spec); spec.loader.exec_module(b200)
def P(*a): print(*a, flush=True)
mode = sys.argv[1]
m, n = int(sys.argv[2]), int(sys.argv[3]); hom = (m // 12, m * 9 // 10)
a, b = synth.make_pair(m, n, [hom], 0.05, 0.01, 0.01, 0, 5)
al = b200.Aligner(kernel=b200.KERNEL_S16X2)
al.set_sequences(a, b)
kw = dict(row=dict(want_last_row=True), col=dict(want_last_column=True), colnocb=dict(want_last_column=True, use_callbacks=False))[mode]
P("start", mode)
t0 = time.time()
r = al.align_partition(prune=True, **kw); P("prune", r["best"], r["cells"], time.time() - t0)
