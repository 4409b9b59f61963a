#!/bin/bash
# Scaling check on N GPUs of one box (gpurun --gpus N): the driver's own launch line for our arm, then for the reference arm.
# usage: gpu_scale.sh N [extra bench args]
N=${1:-8}; shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 \
  bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
echo "ours N=$N rc=$?"; python scripts/show_bench.py gpurun_out/scale_n$N.json 2>/dev/null | head -4
grep -i "error\|Traceback" gpurun_out/scale_n$N.err | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29812 \
  bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/scale_ref_n$N.json 2> gpurun_out/scale_ref_n$N.err
echo "reference N=$N rc=$?"; cut -c1-200 gpurun_out/scale_ref_n$N.json
