#!/bin/bash
# validation of the plane-sourced positive-pair paths, matching kernels, tiled combine; A/B benches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -20
SSP_POS_FWD=nchw SSP_POS_EPI=fp32 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "descriptor or loss_step" > gpurun_out/pytest_old.log 2>&1; echo "pytest(old paths) rc=$? $(tail -1 gpurun_out/pytest_old.log)"
for v in planes epi_fp32 fwd_nchw; do
  case $v in
    planes) export SSP_POS_FWD=planes SSP_POS_EPI=planes;;
    epi_fp32) export SSP_POS_FWD=planes SSP_POS_EPI=fp32;;
    fwd_nchw) export SSP_POS_FWD=nchw SSP_POS_EPI=planes;;
  esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-adapt --no-semantic > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
  python scripts/show_bench.py gpurun_out/bench_$v.json 2>/dev/null | head -16
done
