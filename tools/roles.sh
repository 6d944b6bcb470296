#!/bin/bash
# Needs a tuning build: make -C gan-segmentation_b200/csrc clean && make -C gan-segmentation_b200/csrc TUNING=1
# Role isolation of single conv layers (GSX_DBG: 1 no epilogue, 2 no MMA, 4 no loads): prints ms for
# all / MMA alone (5) / epilogue alone (6) / loads alone (3)
for l in "$@"; do
  line="$l"
  for d in 0 5 6 3; do
    ms=$(GSX_DBG=$d timeout 120 python tools/one_conv.py $l 32 2>/dev/null | awk '{print $3}')
    line="$line  dbg$d=$(printf %.3f $ms)"
  done
  echo "$line"
done
