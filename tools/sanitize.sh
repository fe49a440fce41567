#!/bin/bash
# compute-sanitizer over small cases of both paths (memcheck: out-of-bounds / misaligned accesses; racecheck on the SIMT helpers'
# shared-memory use).  usage: tools/sanitize.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_superpoint_gpu.py -m gpu -q -k "all_64x64 or topk_96x128 or rejects" > gpurun_out/${tag}_memcheck_superpoint.log 2>&1; echo "memcheck superpoint rc=$?"; tail -4 gpurun_out/${tag}_memcheck_superpoint.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tiny_n64_m96 or leaf3 or ragged_query_lengths" > gpurun_out/${tag}_memcheck_matcher.log 2>&1; echo "memcheck matcher rc=$?"; tail -4 gpurun_out/${tag}_memcheck_matcher.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_superpoint_gpu.py -m gpu -q -k "all_64x64" > gpurun_out/${tag}_racecheck_superpoint.log 2>&1; echo "racecheck superpoint rc=$?"; tail -4 gpurun_out/${tag}_racecheck_superpoint.log
