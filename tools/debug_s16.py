import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, synth, oracle_lib as O
spec = importlib.util.spec_from_file_location("masa_cudalign_b200", os.path.join(ROOT, "masa-cudalign_b200", "__init__.py"))
b200 = importlib.util.module_from_spec(spec); spec.loader.exec_module(b200)
np.set_printoptions(linewidth=250)
for (m, n) in [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(20, 30), (40, 100), (700, 900)]:
    a, b = synth.make_pair(m, n, [(m // 5, m * 4 // 5)], 0.05, 0.01, 0.01, 0, 1)
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    r = al.align_partition(want_last_row=True, want_last_column=True)
    o = O.full_matrix(a, b, O.SW, row_ids=[m - 1])
    print("==", m, n, "best", r["best"], o["best"])
    lr, olr = r["rows"][m], o["rows"][m - 1]
    bad = np.nonzero((lr["h"] != olr["h"]) | (lr["x"] != olr["x"]))[0]
    print("last row mismatches:", bad.size, bad[:20])
    if bad.size:
        k = bad[0]; print(" got", lr[max(0,k-2):k+6]); print(" exp", olr[max(0,k-2):k+6])
    lc, olc = r["last_column"], o["last_col"]
    bad = np.nonzero((lc["h"] != olc["h"]) | (lc["x"] != olc["x"]))[0]
    print("last col mismatches:", bad.size, bad[:20])
    if bad.size:
        k = bad[0]; print(" got", lc[max(0,k-2):k+6]); print(" exp", olc[max(0,k-2):k+6])
    al.close()
