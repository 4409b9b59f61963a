#!/bin/bash
# Round-2 call 1: system info + parity/A-B of the default-off variants written at the end of round 1.
mkdir -p gpurun_out
{ nvidia-smi topo -m; nvidia-smi -L; lscpu | head -40; numactl -H 2>&1 | head -20; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c; nproc; free -g; } > gpurun_out/sysinfo.txt 2>&1
bash scripts/gpu_r2_ab.sh
