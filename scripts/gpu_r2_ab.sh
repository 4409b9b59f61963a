#!/bin/bash
# Round-2 opener: parity + A/B of the three default-off kernel variants written at the end of round 1
# (SSP_FWD_EPI=2, SSP_BG_BITS=deep, SSP_BG_POS=late).  Each variant: descriptor/loss-step parity tests, then a bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; print(g.build())" > gpurun_out/build.log 2>&1; echo "build rc=$?"
run() { # tag, env assignments...
  local tag=$1; shift
  env "$@" timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "descriptor or loss_step" > gpurun_out/t_$tag.log 2>&1
  echo "== $tag tests rc=$? $(tail -1 gpurun_out/t_$tag.log)"
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-adapt --no-semantic > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "   $tag bench rc=$?"; python scripts/show_bench.py gpurun_out/bench_$tag.json 2>/dev/null | head -6
}
run base SSP_NONE=1
run fwdepi2 SSP_FWD_EPI=2
run deepbits SSP_BG_BITS=deep
run latepos SSP_BG_POS=late
run all3 SSP_FWD_EPI=2 SSP_BG_BITS=deep SSP_BG_POS=late
run fold SSP_FWD_EPI=2 SSP_BG_ALPHA=fold
run all4 SSP_FWD_EPI=2 SSP_BG_BITS=deep SSP_BG_POS=late SSP_BG_ALPHA=fold
