#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU suite (memcheck on everything, racecheck on the non-TMA kernels)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
K='small_golden or test_nms_golden or test_detector_loss_pair or test_flatten_combine or test_valid_mask or test_warp_points or test_box_nms or test_inv_warp_golden'
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_memcheck.log | tr '\n' ' ')"
K2='test_nms_golden or test_detector_loss_pair or test_flatten_combine or test_valid_mask or test_warp_points or test_box_nms'
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "$K2" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_racecheck.log | tr '\n' ' ')"
