#!/bin/bash
# Round 2, GPU call 5 (one GPU): whole GPU suite on the pipelined-launch build, default bench line, 4096^2 line.
out=gpurun_out/r2_call5
mkdir -p $out
timeout -k 5 1800 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -60 | tee $out/tests.log
timeout -k 5 400 python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 400 $out/bench_default.json
timeout -k 5 200 python bench.py --grid-n 4096 --no-extras --no-cpu-baseline > $out/bench_4096.json 2> $out/bench_4096.err; head -c 300 $out/bench_4096.json
timeout -k 5 200 python bench.py --grid-n 128 --iters 352 --steps 20 --no-extras --no-cpu-baseline > $out/bench_128.json 2> $out/bench_128.err; head -c 300 $out/bench_128.json
