#!/bin/bash
# Full SW alignment (stages 1-6) of the BASELINE cfg2 pair (5M x 5M synthetic, --ram-size=8G) with build/cudalign:
# process wall time, MASA-Core's per-stage split, stage-1 crosspoint and a checksum of alignment.00.bin.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d /tmp/cfg2full.XXXXXX); cd "$W"
python "$ROOT/tools/synth.py" --config cfg2 --scale ${1:-1.0} --out s > /dev/null
T0=$(date +%s.%N)
{ time B200_DEBUG=${B200_DEBUG_FLAG:-} "$ROOT/build/cudalign" --work-dir=w --clear --verbose=0 --ram-size=8G s_A.fa s_B.fa > log.txt 2>&1 ; } 2> time.txt || { tail -20 log.txt; exit 1; }
T1=$(date +%s.%N)
tr '\n' ' ' < time.txt; echo
python3 -c "print('process wall %.2f s' % ($T1 - $T0))"
grep -E "SEQUENCES|INIT|STAGE[1-6]|TOTAL" w/statistics | tr -s ' ' | tr '\n' '|'; echo
tr '\n' ' ' < w/crosspoints/crosspoint_01.00; echo
sha256sum w/alignment.00.bin | cut -c1-16; ls -la w/alignment.00.bin | awk '{print $5 " bytes"}'
grep -h "B200 partitions" w/statistics.ALIGNER | head -2
grep "sra_prefault" log.txt
rm -rf "$W"
