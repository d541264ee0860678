#!/bin/bash
# Round 2, GPU call 6 (one GPU): whole GPU suite on the build with idle-warp skipping, default bench line, ncu launch list.
out=gpurun_out/r2_call6
mkdir -p $out
timeout -k 5 1800 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -60 | tee $out/tests.log
timeout -k 5 400 python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 300 $out/bench_default.json
cmd="python bench.py --steps 2 --warmup 1 --iters 40 --no-cpu-baseline --no-e2e --no-extras"
$cmd > $out/plain.log 2>&1 &&
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv $cmd > $out/ncu.log 2>&1
tail -2 $out/ncu.log
