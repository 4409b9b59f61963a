#!/bin/bash
# N-GPU bench under torchrun (weak scaling, global normalisers over NCCL inside the CUDA graph)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-graph --no-adapt > gpurun_out/bench_n${N}_eager.json 2> gpurun_out/bench_n${N}_eager.err; echo "n$N eager rc=$?"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref n$N rc=$?"
grep -v Warning gpurun_out/bench_n$N.err | tail -n 5
