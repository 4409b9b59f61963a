#!/bin/bash
# Round-2 call 2: timeline trace of the tcgen05 kernels + isolation timings
mkdir -p gpurun_out
SSP_TRACE=1 timeout 300 python scripts/trace_desc.py gpurun_out/trace.npz > gpurun_out/trace.log 2>&1; echo "trace rc=$?"; tail -3 gpurun_out/trace.log
timeout 300 python scripts/micro_desc.py > gpurun_out/micro.log 2>&1; echo "micro rc=$?"; cat gpurun_out/micro.log
