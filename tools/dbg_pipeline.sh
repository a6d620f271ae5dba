cd /tmp && rm -rf dbg && mkdir dbg && cd dbg
python /root/repo/tools/synth.py --m 3000 --n 2700 --a0 500 --a1 2500 --ps 0.05 --pd 0.02 --pi 0.02 --seed 11 --out s >/dev/null
for args in "--no-fast-path --no-block-pruning --stage-1" "--no-fast-path --stage-1" "--no-fast-path --kernel=s32 --stage-1" "--no-fast-path --kernel=s32 --no-block-pruning --stage-1" "--no-fast-path --blocks=1 --stage-1"; do
  /root/repo/build/cudalign --work-dir=wd --clear --verbose=0 $args s_A.fa s_B.fa > log 2>&1; echo "$args rc=$? $(cat wd/crosspoints/crosspoint_01.00 | tr '\n' ' ')"
done
