#!/bin/bash
# compute-sanitizer over the one-kernel ego stage (named barriers per warp group, per-warp shared-memory scratch, staging tiles):
# the small cases of tests/test_fusion_gpu.py::test_ego_att_one_kernel.  usage: tools/sanitize_ego.sh [outdir]
out=${1:-gpurun_out}
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_fusion_gpu.py -q -x -k "test_ego_att_one_kernel and not 100x352" \
    > $out/r2_sanitizer_ego_att_$tool.log 2>&1
  echo "rc=$?" >> $out/r2_sanitizer_ego_att_$tool.log
  tail -3 $out/r2_sanitizer_ego_att_$tool.log
done
