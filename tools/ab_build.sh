#!/bin/bash
# A/B of a compile-time macro on the GPU box: tools/ab_build.sh MACRO "v1 v2 ..." [shape]   (rebuilds the library per value)
MACRO=$1; VALS=$2; SHAPE=${3:-5w20s}
for v in $VALS; do
  (cd meta-fine-tuning_b200/csrc && rm -f *.o && make -j EXTRA=-D$MACRO=$v > /dev/null 2>&1) || { echo "build failed for $v"; continue; }
  echo "$MACRO=$v"; bash tools/r02_ab.sh MFT_BWD_SPLIT "59 59" $SHAPE 2>&1 | cut -c1-60
done
(cd meta-fine-tuning_b200/csrc && rm -f *.o && make -j > /dev/null 2>&1)
