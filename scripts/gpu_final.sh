#!/bin/bash
# what the driver runs at round end, in the same order
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench default rc=$? lines=$(wc -l < gpurun_out/final_bench_default.json)"
python scripts/show_bench.py gpurun_out/final_bench_default.json | head -30
cat gpurun_out/final_ref.json | cut -c1-400
