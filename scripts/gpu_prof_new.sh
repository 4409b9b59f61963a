#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1
timeout 100 ncu --set full --clock-control none --import-source on \
  -k regex:"sem_ce_up8|desc_pos_fwd_planes|desc_bits_gemm_tc|combine_heatmap_kernel|nms_round_square|desc_pos_coef|sem_scale" -s 7 -c 12 \
  -o gpurun_out/prof_new python scripts/prof_new.py > gpurun_out/ncu_new.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_new.log
