#!/bin/bash
# Round 2, GPU call 11 (one GPU): whole GPU suite + default bench on the build with the size-dependent strip geometry.
out=gpurun_out/r2_call11
mkdir -p $out
timeout -k 5 1800 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -40 | tee $out/tests.log
timeout -k 5 400 python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 300 $out/bench_default.json
timeout -k 5 200 python bench.py --grid-n 4096 --no-extras --no-cpu-baseline > $out/bench_4096.json 2> $out/bench_4096.err; head -c 200 $out/bench_4096.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
