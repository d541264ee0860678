#!/bin/bash
# Round 2, GPU call 12 (one GPU): the whole GPU suite + smoke + default bench on the final build.
out=gpurun_out/r2_call12
mkdir -p $out
timeout -k 5 1800 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -30 | tee $out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/smoke.log
timeout -k 5 400 python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 300 $out/bench_default.json
