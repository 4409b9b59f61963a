#!/bin/bash
# 2-GPU bench under torchrun exactly as the driver launches it (weak scaling, NCCL exchange inside the CUDA graph)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1
N=${1:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"
python scripts/show_bench.py gpurun_out/bench_n$N.json | head -8
grep -v Warning gpurun_out/bench_n$N.err | tail -n 4
