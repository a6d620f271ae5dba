#!/bin/bash
# Full SW alignment (stages 1-6) of the cfg2-shaped 5M x 5M synthetic pair with build/cudalign: wall time per stage.
set -e
SCALE=${1:-1.0}
W=/tmp/cfg2 && rm -rf $W && mkdir -p $W && cd $W
python /root/repo/tools/synth.py --config cfg2 --scale $SCALE --out s > /dev/null
T0=$(date +%s.%N)
B200_DEBUG=${B200_DEBUG_FLAG:-} /root/repo/build/cudalign --work-dir=w --clear --verbose=0 --ram-size=8G s_A.fa s_B.fa > log.txt 2>&1 || { tail -20 log.txt; exit 1; }
T1=$(date +%s.%N)
python3 -c "print(\"wall_total_s\", $T1 - $T0)"
for s in 1 2 3 4 5 6; do f=w/statistics_0$s.00; [ -f $f ] && grep -E "^ *Total|TOTAL|Time|time" $f | tail -2 | sed "s/^/stage$s: /"; done
cat w/crosspoints/crosspoint_01.00 | tr '\n' ' '; echo
head -c 600 w/statistics | tail -c 400
ls -la w/alignment.00.bin; grep -h "B200 partitions" w/statistics* | head -3
