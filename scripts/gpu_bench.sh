#!/bin/bash
# bench + ncu evidence on one B200.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --engine bf16 --no-adapt --no-cpu > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --engine fp32 --no-adapt --no-cpu > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"desc_dense_fwd_tc|desc_bits_gemm_tc" -s 4 -c 3 \
  -o gpurun_out/prof_desc python bench.py --steps 2 --warmup 3 --no-adapt --no-cpu > gpurun_out/ncu_desc.log 2>&1; echo "ncu desc rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"combine_heatmap|detector_loss_fwd|desc_pack|flatten_detection" -s 4 -c 4 \
  -o gpurun_out/prof_hbm python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
ls -la gpurun_out
