#!/usr/bin/env python3
"""Synthetic DNA pairs with planted, mutated homology (NCBI data is not available offline).

seqA = i.i.d. uniform ACGT of length m.  seqB = flankL | mutate(seqA[a0:a1]) | flankR, padded with random
bases / truncated to length n.  mutate(): per-base substitution with prob p_s (uniform over the 3 other
bases), deletion with prob p_d, insertion after the base with prob p_i of a geometric(0.5)-length random
string; optionally K large indels of length U[100,5000].  Deterministic in (seed, parameters); the RNG is
numpy's MT19937 bit generator.  Used by tests/, bench.py and tools/; never by the product path.
"""
import argparse
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

CONFIGS = {
    # name: (m, n, segments [(a0, a1, shift_in_B)], p_s, p_d, p_i, K, seed)       -- SURVEY.md 8d rows cfg1..cfg5
    "cfg1": dict(m=1_000_000, n=1_000_000, segs=[(100_000, 900_000)], p_s=0.05, p_d=0.01, p_i=0.01, K=0, seed=0xC0DA0001),
    "cfg2": dict(m=5_000_000, n=5_000_000, segs=[(500_000, 2_400_000), (2_600_000, 4_500_000)], p_s=0.03, p_d=0.004, p_i=0.004, K=20, seed=0xC0DA0002, shift=60_000),
    "cfg3": dict(m=23_000_000, n=25_000_000, segs=[(1_000_000, 22_000_000)], p_s=0.10, p_d=0.025, p_i=0.025, K=200, seed=0xC0DA0003),
    "cfg4": dict(m=48_000_000, n=46_000_000, segs=[(0, 48_000_000)], p_s=0.012, p_d=0.0015, p_i=0.0015, K=500, seed=0xC0DA0004),
    "cfg5": dict(m=249_000_000, n=228_000_000, segs=[(0, 120_000_000), (140_000_000, 249_000_000)], p_s=0.012, p_d=0.0015, p_i=0.0015, K=3000, seed=0xC0DA0005),
}


def _rand_bases(rng, k):
    return ACGT[rng.integers(0, 4, size=k, dtype=np.uint8)]


def mutate(rng, seg, p_s, p_d, p_i, K=0):
    """Return a mutated copy of the uint8 base array `seg`."""
    L = seg.size
    out = seg.copy()
    if L == 0:
        return out
    # substitutions: add 1..3 (mod 4) to the base index
    idx = np.searchsorted(ACGT, out)               # A,C,G,T are sorted in ASCII
    sub = rng.random(L) < p_s
    idx = np.where(sub, (idx + rng.integers(1, 4, size=L)) % 4, idx)
    out = ACGT[idx]
    keep = rng.random(L) >= p_d
    ins = rng.random(L) < p_i
    ins_len = np.where(ins, rng.geometric(0.5, size=L), 0)
    # K large indels: half deletions (clear `keep` over a run), half insertions (one long insert)
    for k in range(K):
        pos = int(rng.integers(0, L))
        ln = int(rng.integers(100, 5001))
        if k % 2 == 0:
            keep[pos:pos + ln] = False
        else:
            ins_len[pos] += ln
    reps = keep.astype(np.int64) + ins_len
    total = int(reps.sum())
    res = np.empty(total, dtype=np.uint8)
    # position of each original base's block in the output
    starts = np.cumsum(reps) - reps
    kept_pos = starts[keep]
    res[:] = _rand_bases(rng, total)               # inserted material (random); kept bases overwrite below
    res[kept_pos] = out[keep]
    return res


def make_pair(m, n, segs, p_s, p_d, p_i, K=0, seed=1, shift=0):
    """Return (seqA, seqB) as uint8 numpy arrays of ASCII A/C/G/T with lengths exactly (m, n)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    a = _rand_bases(rng, m)
    parts = []
    prev_end = 0
    bpos = 0
    for si, (a0, a1) in enumerate(segs):
        a0 = max(0, min(a0, m)); a1 = max(a0, min(a1, m))
        # non-homologous flank of the same length as the gap in A (plus `shift` extra before later segments)
        gap = a0 - prev_end + (shift if si > 0 else 0)
        parts.append(_rand_bases(rng, max(gap, 0)))
        parts.append(mutate(rng, a[a0:a1], p_s, p_d, p_i, K if si == 0 else 0))
        prev_end = a1
    parts.append(_rand_bases(rng, max(m - prev_end, 0)))
    b = np.concatenate(parts) if parts else np.empty(0, np.uint8)
    if b.size < n:
        b = np.concatenate([b, _rand_bases(rng, n - b.size)])
    return a, b[:n].copy()


def make_config(name, scale=1.0):
    """Pair for a named BASELINE config, optionally scaled down (all lengths multiplied by `scale`)."""
    c = dict(CONFIGS[name])
    m = max(1, int(c["m"] * scale)); n = max(1, int(c["n"] * scale))
    segs = [(int(a0 * scale), int(a1 * scale)) for a0, a1 in c["segs"]]
    K = int(round(c["K"] * min(1.0, scale))) if c["K"] else 0
    return make_pair(m, n, segs, c["p_s"], c["p_d"], c["p_i"], K, c["seed"], int(c.get("shift", 0) * scale))


def write_fasta(path, seq, header):
    with open(path, "wb") as f:
        f.write(b">" + header.encode() + b"\n")
        L = seq.size
        full = (L // 70) * 70
        if full:
            block = np.empty((full // 70, 71), dtype=np.uint8)
            block[:, :70] = seq[:full].reshape(-1, 70)
            block[:, 70] = 10
            f.write(block.tobytes())
        if L > full:
            f.write(seq[full:].tobytes() + b"\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--m", type=int); ap.add_argument("--n", type=int)
    ap.add_argument("--a0", type=int, default=0); ap.add_argument("--a1", type=int, default=0)
    ap.add_argument("--ps", type=float, default=0.05); ap.add_argument("--pd", type=float, default=0.01)
    ap.add_argument("--pi", type=float, default=0.01); ap.add_argument("--K", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", required=True, help="output prefix: writes <out>_A.fa and <out>_B.fa")
    args = ap.parse_args()
    if args.config:
        a, b = make_config(args.config, args.scale)
        tag = f"{args.config}x{args.scale}"
    else:
        a, b = make_pair(args.m, args.n, [(args.a0, args.a1)], args.ps, args.pd, args.pi, args.K, args.seed)
        tag = f"custom seed={args.seed}"
    write_fasta(args.out + "_A.fa", a, f"synth_{tag}_A")
    write_fasta(args.out + "_B.fa", b, f"synth_{tag}_B")
    print(a.size, b.size)
