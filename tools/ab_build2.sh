#!/bin/bash
# A/B of compile-time macros on the GPU box: tools/ab_build2.sh "FLAGS1" "FLAGS2" ...   (each a full -D list)
for f in "$@"; do
  (cd meta-fine-tuning_b200/csrc && rm -f *.o && make -j EXTRA="$f" > /dev/null 2>&1) || { echo "build failed for $f"; continue; }
  echo "== $f"; grep -A2 "score8_kernelILb1ELb0ELi3\|dy8_kernelILb1ELb0ELi3" meta-fine-tuning_b200/csrc/wcompute.ptxas.log | grep -E "Used|spill" | head -4
  bash tools/r02_ab.sh MFT_BWD_SPLIT "59 59" 2>&1 | cut -c1-60
done
(cd meta-fine-tuning_b200/csrc && rm -f *.o && make -j > /dev/null 2>&1)
