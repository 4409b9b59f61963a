#!/bin/bash
# N-GPU repetition harness: RP runs with the peer-memory exchange, RA with the all-reduce fallback (rc and line per run)
# usage: gpu_r2_reps.sh N RP RA
N=${1:-2}; RP=${2:-9}; RA=${3:-2}
mkdir -p gpurun_out
COMMON="bench.py --gpus $N --steps 40 --warmup 3 --no-cpu --no-adapt --no-semantic --no-variants"
run() { local tag=$1 port=$2; shift 2
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port $COMMON \
    > gpurun_out/rep_$tag.json 2> gpurun_out/rep_$tag.err
  echo "$tag rc=$? incomplete=$(grep -c incomplete gpurun_out/rep_$tag.json) $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/rep_$tag.json').read().strip().splitlines()[-1]); print('value=%.0f ms=%.3f e2e=%s'%(d['value'],d['ms_per_step'],d['e2e'].get('value')))
except Exception as e: print('no line', e)")"
}
for i in $(seq 1 $RP); do run b_p2p_$i $((29600+i)) SSP_X=1; done
for i in $(seq 1 $RA); do run b_ar_$i $((29700+i)) SSP_EXCHANGE=allreduce; done
