#!/bin/bash
# the igemm launches of the frame by index inside the NVTX range "timed": 0-18 convs, 19-21 deblocks, 22 / 23 shrinker, 24 encode
out=${1:-gpurun_out}
for spec in "shrink0:22" "shrink1:23" "encode:24" "deblock2:21" "s2conv:12"; do
  name=${spec%%:*}; idx=${spec#*:}
  ncu --set full --clock-control none --nvtx --nvtx-include "timed/" -k regex:igemm_kernel -s $idx -c 1 -o $out/r2_$name -f \
      python tools/frame_once.py 8 1 > $out/ncu_$name.log 2>&1
  grep -c "==PROF== Profiling" $out/ncu_$name.log
done
ls -la $out/*.ncu-rep
