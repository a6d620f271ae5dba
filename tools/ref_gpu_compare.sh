#!/bin/bash
# Kernel-vs-kernel baseline (SURVEY.md 8d "Reference GPU baseline", optional): stage 1 of the same FASTA pair through
#   (a) oracle/_ref/cudalign_ref_gpu -- the reference's own CUDA aligner (R/src/CUDAligner.cu kernels, grid policy and host
#       loop unmodified; texture references replaced by __ldg pointers so that it compiles with CUDA 12, oracle/build_ref_gpu.sh)
#   (b) build/cudalign              -- this repo's aligner behind the same MASA-Core driver
# with the same flags (--stage-1 --no-flush: best score + position only), pruning on and off.  Prints the ALIGN time of
# MASA-Core's own statistics_01.00, GCUPS = m*n / that time, and the stage-1 crosspoint of both (must be identical).
#   usage: tools/ref_gpu_compare.sh <cfg> <scale> [per-run timeout in s]
set -u
CFG=${1:-cfg1}; SCALE=${2:-1.0}; TMO=${3:-200}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d /tmp/refgpu_cmp.XXXXXX)
cd "$W"
python "$ROOT/tools/synth.py" --config "$CFG" --scale "$SCALE" --out s > dims.txt
read M N < dims.txt
echo "pair: $CFG x $SCALE = $M x $N"
run() {   # label binary extra-flags...
  local label=$1 exe=$2; shift 2
  rm -rf "w_$label"
  local t0=$(date +%s.%N)
  timeout "$TMO" "$exe" --work-dir="w_$label" --clear --verbose=0 --stage-1 --no-flush "$@" s_A.fa s_B.fa > "log_$label.txt" 2>&1
  local rc=$?
  local t1=$(date +%s.%N)
  if [ $rc -ne 0 ]; then echo "$label: rc=$rc (timeout ${TMO}s or failure)"; tail -3 "log_$label.txt"; return; fi
  python3 - "$label" "$M" "$N" "w_$label" "$t0" "$t1" <<'PY'
import sys, re
label, m, n, wd, t0, t1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], float(sys.argv[5]), float(sys.argv[6])
st = open(wd + "/statistics_01.00").read()
align = float(re.search(r"ALIGN:\s+([0-9.]+)", st).group(1))
cells = re.search(r"Cells:\s+(\S+)\s+\((\S+)\)", st)
xp = open(wd + "/crosspoints/crosspoint_01.00").read().split("\n")[1]
print(f"{label:28s} ALIGN {align:10.1f} ms  {m * n / align / 1e6:9.1f} GCUPS (m*n/t)  cells {cells.group(1)} ({cells.group(2)})  crosspoint_01 {xp}  process wall {t1 - t0:.1f} s")
PY
}
REF_EXE=${REF_EXE:-$ROOT/oracle/_ref/cudalign_ref_gpu}; NEW_EXE=${NEW_EXE:-$ROOT/build/cudalign}
run ref_gpu_pruning   "$REF_EXE"
run b200_pruning      "$NEW_EXE"
run ref_gpu_nopruning "$REF_EXE" --no-block-pruning
run b200_nopruning    "$NEW_EXE" --no-block-pruning
rm -rf "$W"
