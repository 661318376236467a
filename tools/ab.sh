#!/bin/bash
# A/B of several builds of the library on the same box: tools/ab.sh [-r ROUNDS] libA.so libB.so ...
# (alternating rounds of c2, c3, c1 through tools/prof_run.py)
rounds=3
if [ "$1" = "-r" ]; then rounds=$2; shift 2; fi
for round in $(seq $rounds); do
  for v in "$@"; do
    for w in c2 c3 c1; do
      RTP_B200_LIB=$PWD/$v python tools/prof_run.py $w 8 2>&1 | head -1 | sed "s|^|$(basename $v) |"
    done
  done
done
