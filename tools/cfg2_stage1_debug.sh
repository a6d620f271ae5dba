#!/bin/bash
# Where the wall time of stage 1 goes in the full cfg2 alignment (5M x 5M, --ram-size=8G): MASA-Core's own stage statistics
# plus the engine's B200_DEBUG timestamps (setup / kernel / special rows streamed / dispatch).
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d /tmp/cfg2dbg.XXXXXX); cd "$W"
python "$ROOT/tools/synth.py" --config cfg2 --scale ${1:-1.0} --out s > /dev/null
B200_DEBUG=1 "$ROOT/build/cudalign" --work-dir=w --clear --verbose=0 --ram-size=8G --stage-1 s_A.fa s_B.fa > log.txt 2>&1
grep "b200\]" log.txt | head -20
sed -n '/Stage1 times/,/MCUPS/p' w/statistics_01.00
grep -E "SEQUENCES|INIT|STAGE1|TOTAL" w/statistics
cat w/statistics.ALIGNER 2>/dev/null | head -20
rm -rf "$W"
