#!/bin/bash
# A/B of two builds of the library on the same box: tools/ab.sh libA.so libB.so  (three alternating rounds of c2, c3, c1)
for round in 1 2 3; do
  for v in "$1" "$2"; do
    for w in c2 c3 c1; do
      RTP_B200_LIB=$PWD/$v python tools/prof_run.py $w 8 2>&1 | head -1 | sed "s|^|$(basename $v) |"
    done
  done
done
