#!/bin/bash
# compute-sanitizer memcheck over the small-size GPU parity tests (every kernel family through the C ABI)
mkdir -p gpurun_out
K=${1:-"golden or identity_kat or warp_points or labels_and_masks or test_detector_loss or box_nms or abi_errors or other_channel or exchange"}
timeout 1400 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_tests.log \
  python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "$K" > gpurun_out/memcheck_tests.out 2>&1
echo "memcheck rc=$? $(tail -1 gpurun_out/memcheck_tests.out)"; tail -3 gpurun_out/memcheck_tests.log
