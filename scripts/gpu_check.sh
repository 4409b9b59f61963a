#!/bin/bash
# Runs the GPU parity suite in separate processes (a hung kernel only costs its own slice).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1
run() { # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  timeout $to python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1
  echo "== $name exit=$? $(tail -1 gpurun_out/$name.log)"
}
run desc_bf16x3 300 -k "descriptor and bf16x3"
run desc_bf16 300 -k "descriptor and bf16 and not bf16x3"
run desc_fp32 300 -k "descriptor and fp32 or other_channel"
run desc_rest 600 -k "engines_agree or kitti or loss_step"
run simple 600 -k "not descriptor and not loss_step"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit=$? $(tail -1 gpurun_out/smoke.log)"
