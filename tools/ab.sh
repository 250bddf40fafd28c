#!/bin/bash
# A/B of two library builds on ONE box, alternating: tools/ab.sh <libA.so> <libB.so> [rounds]   (env: MODEL, BATCH, FLAGS)
a=$1; b=$2; n=${3:-3}
for i in $(seq $n); do
  VITB200_LIB=$a python tools/time_forward.py
  VITB200_LIB=$b python tools/time_forward.py
done
