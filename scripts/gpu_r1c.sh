#!/bin/bash
# full parity suite on the reworked kernels + A/B of the bits-GEMM ring depth / item schedule
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -20
for v in base deep rr deeprr; do
  case $v in
    base) export SSP_BG_NS=x SSP_BG_SCHED=x;;
    deep) export SSP_BG_NS=deep SSP_BG_SCHED=x;;
    rr) export SSP_BG_NS=x SSP_BG_SCHED=rr;;
    deeprr) export SSP_BG_NS=deep SSP_BG_SCHED=rr;;
  esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-adapt --no-semantic > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
  python scripts/show_bench.py gpurun_out/bench_$v.json 2>/dev/null | head -16
done
