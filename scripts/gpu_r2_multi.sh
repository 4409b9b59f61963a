#!/bin/bash
# Round-2 multi-GPU check (gpurun --gpus 2): exchange-kernel tests, repeated 2-GPU bench runs on both exchange backends
# (p2p kernel = product path; allreduce = one captured NCCL all-reduce, the configuration closest to round 1's failure),
# then compute-sanitizer memcheck of the p2p path under torchrun.   usage: gpu_r2_multi.sh [reps_p2p] [reps_allreduce]
mkdir -p gpurun_out
RP=${1:-8}; RA=${2:-6}; N=2
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/t_multi.log 2>&1; echo "multi tests rc=$? $(tail -1 gpurun_out/t_multi.log)"
COMMON="bench.py --gpus $N --steps 40 --warmup 3 --no-cpu --no-adapt --no-semantic"
run() { # tag port env...
  local tag=$1 port=$2; shift 2
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port $COMMON \
    > gpurun_out/rep_$tag.json 2> gpurun_out/rep_$tag.err
  echo "$tag rc=$? incomplete=$(grep -c incomplete gpurun_out/rep_$tag.json) $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/rep_$tag.json').read().strip().splitlines()[-1]); print('value=%.0f ms=%.3f e2e=%s'%(d['value'],d['ms_per_step'],d['e2e'].get('value')))
except Exception as e: print('no line', e)")"
}
for i in $(seq 1 $RP); do run p2p_$i $((29600+i)) SSP_X=1; done
for i in $(seq 1 $RA); do run ar_$i $((29700+i)) SSP_EXCHANGE=allreduce; done
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 --no-python \
  compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_n2_%p.log python bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-adapt --no-semantic \
  > gpurun_out/san_mem.json 2> gpurun_out/san_mem.err; echo "memcheck rc=$?"; tail -n 2 gpurun_out/memcheck_n2_*.log
