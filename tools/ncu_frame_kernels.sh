#!/bin/bash
# `ncu --set full` captures of one launch of each kernel family of the 8-agent frame (tools/frame_once.py), for
# profiles/r2_ncu_*.txt (tools/ncu_summary.py).  usage: tools/ncu_frame_kernels.sh [outdir]
out=${1:-gpurun_out}
for spec in "pillar:pillar_bev_kernel" "decode:codebook_decode_kernel" "fuse:fuse_kernel" "heads:heads_kernel" \
            "encode:EncodeEpilogue" "shrink0:FixedEpilogue<3, 0" "shrink1:igemm_kernel<256, 128, 1, FixedEpilogueC<1, 1>, 1, 1, 0" ; do
  name=${spec%%:*}; pat=${spec#*:}
  ncu --set full --clock-control none --nvtx --nvtx-include "timed/" -k "regex:$pat" -c 1 -o $out/r2_$name -f \
      python tools/frame_once.py 8 1 > $out/ncu_$name.log 2>&1
  tail -1 $out/ncu_$name.log
done
ls -la $out/*.ncu-rep
