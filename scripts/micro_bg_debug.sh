#!/bin/bash
# which feed limits the indicator GEMM: run the isolation timings with the MMA thread ignoring the A ring (1), the B ring (2), both (3)
for d in 0 1 2 3; do echo "== SSP_BG_DEBUG=$d"; SSP_BG_DEBUG=$d timeout 200 python scripts/micro_desc.py 2>&1 | grep "bwd"; done
