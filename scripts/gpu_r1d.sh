#!/bin/bash
# parity suite + bench + ncu evidence for the reworked kernels (detector, pos_fwd, bits GEMM schedule, tiled combine, semantic)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_v6.json 2>/dev/null | head -24
SSP_COMBINE=gather timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-semantic > gpurun_out/bench_gather.json 2> gpurun_out/bench_gather.err; echo "bench gather rc=$?"
python scripts/show_bench.py gpurun_out/bench_gather.json 2>/dev/null | grep adapt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_v6.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"sem_ce_up8|detector_loss_fwd|detector_loss_bwd|desc_pos_fwd|desc_bits_gemm|desc_dense_fwd_tc|combine_heatmap_tiled|desc_pos_coef|desc_finalize|nms_round_square" -s 20 -c 24 \
  -o gpurun_out/prof_v6 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -30
