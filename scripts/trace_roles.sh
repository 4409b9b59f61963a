#!/bin/bash
mkdir -p gpurun_out
for r in 1 2 4 8; do
  SSP_TRACE=1 SSP_TRACE_ROLES=$r timeout 200 python scripts/trace_desc.py gpurun_out/trace_r$r.npz > gpurun_out/trace_r$r.log 2>&1; echo "roles=$r rc=$? $(tail -2 gpurun_out/trace_r$r.log | tr '\n' ' ')"
done
