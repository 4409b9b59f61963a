#!/bin/bash
# Parameterised GPU-lease runner (replaces the one-off gpu_r1*.sh scripts).  Every step runs under its own `timeout`
# so that a hung kernel cannot hold the box.  usage: bash scripts/gpu_run.sh step [step ...]   steps:
#   tests[:EXPR]   pytest -m gpu (optionally -k EXPR)          micro     isolation timings of the tcgen05 kernels
#   bench[:ARGS]   bench.py (1 GPU) with extra ARGS             trace     timeline trace (SSP_TRACE build)
#   smoke          __graft_entry__.smoke()                      ref       bench.py --impl reference
#   launches       ncu launch list of bench.py --no-graph       ncufull:REGEX  ncu --set full of the matching kernels
mkdir -p gpurun_out
for step in "$@"; do
  name=${step%%:*}; arg=""; [[ "$step" == *:* ]] && arg=${step#*:}
  case $name in
    tests) if [ -n "$arg" ]; then timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "$arg" > gpurun_out/t_gpu.log 2>&1; else timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; fi
           echo "tests[$arg] rc=$? $(tail -1 gpurun_out/t_gpu.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/t_gpu.log | head -20 ;;
    micro) timeout 300 python scripts/micro_desc.py > gpurun_out/micro.log 2>&1; echo "micro rc=$?"; cat gpurun_out/micro.log | tail -20 ;;
    trace) SSP_TRACE=1 timeout 300 python scripts/trace_desc.py gpurun_out/trace.npz > gpurun_out/trace.log 2>&1; echo "trace rc=$?"; tail -3 gpurun_out/trace.log ;;
    bench) timeout 600 python bench.py $arg > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/bench.json 2>/dev/null | head -24 ;;
    ref)   timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-300 ;;
    launches) # per-launch device times of two graph-free steps (ncu serialises kernels: compare SHARES, not absolutes)
           timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
             python bench.py --steps 2 --warmup 1 --no-cpu --no-adapt --no-semantic --no-graph > gpurun_out/launches.log 2>&1
           echo "launches rc=$?"; python scripts/launch_table.py gpurun_out/launches.csv | head -40 ;;
    launches_graph) # the kernels of the REPLAYED graph, caches left warm between kernels as in the real step (still serialised by ncu)
           timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --graph-profiling node -c 500 --csv --log-file gpurun_out/launches_graph.csv \
             python bench.py --steps 3 --warmup 3 --no-cpu --no-adapt --no-semantic --no-variants > gpurun_out/launches_graph.log 2>&1
           echo "launches_graph rc=$?"; python scripts/launch_table.py gpurun_out/launches_graph.csv | head -40 ;;
    launches_adapt) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_adapt.csv \
             python scripts/prof_adapt.py > gpurun_out/launches_adapt.log 2>&1
           echo "launches_adapt rc=$?"; python scripts/launch_table.py gpurun_out/launches_adapt.csv | head -30 ;;
    ncufull_adapt) # --set full of the warp / aggregation kernels matching regex $arg (scripts/prof_adapt.py, 4 source images)
           ADAPT_I=4 timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$arg" -c 12 -f -o gpurun_out/prof_adapt \
             python scripts/prof_adapt.py > gpurun_out/ncufull_adapt.log 2>&1
           echo "ncufull_adapt rc=$?"; ls -la gpurun_out/prof_adapt.ncu-rep ;;
    ncufull) # --set full capture of the kernels matching regex $arg
           timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$arg" -s 4 -c 6 -f -o gpurun_out/prof_full \
             python bench.py --steps 2 --warmup 1 --no-cpu --no-adapt --no-semantic --no-graph > gpurun_out/ncufull.log 2>&1
           echo "ncufull rc=$?"; ls -la gpurun_out/prof_full.ncu-rep ;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log ;;
    *) echo "unknown step $step" ;;
  esac
done
