#!/bin/bash
# compute-sanitizer over the hand-rolled synchronisation protocols (mbarrier rings, TMEM slots, cp.async side inputs,
# peer stores): small cases of the layer / codebook / fusion tests under memcheck, racecheck and synccheck.
# usage: tools/sanitize.sh [outdir]      (logs: <outdir>/r2_sanitizer_<tool>.log)
out=${1:-gpurun_out}
sel='s0_64_64 or s1_first_64_128_s2 or shrink0_cat384_256 or w4_128_128 or odd_size_64_64 or de1_128_128_s2 or c256_m1_k128_ragged or push_planes or test_fuse'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_layer_gpu.py tests/test_codebook_gpu.py tests/test_fusion_gpu.py -q -x -k "$sel" \
    > $out/r2_sanitizer_$tool.log 2>&1
  echo "rc=$?" >> $out/r2_sanitizer_$tool.log
  tail -4 $out/r2_sanitizer_$tool.log
done
