#!/bin/bash
# focused validation: detector rewrite + semantic kernels + step, then a bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "detector or labels or sem or loss_step" > gpurun_out/t_new.log 2>&1
echo "== new exit=$? $(tail -1 gpurun_out/t_new.log)"
grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/t_new.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_r1b.json 2>/dev/null | head -40
