#!/bin/bash
# Round 2, GPU call 3 (one GPU): the whole GPU test suite on the current build, the default bench line with its extras,
# configs[3] (assembled CSR at 8192^2) as its own line, launch list of the default run, ncu of the CSR kernels.
out=gpurun_out/r2_call3
mkdir -p $out
timeout -k 5 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee $out/tests.log
timeout -k 5 400 python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 600 $out/bench_default.json
timeout -k 5 300 python bench.py --op csr --grid-n 8192 --steps 3 --warmup 3 > $out/bench_csr_8192.json 2> $out/bench_csr.err; tail -c 900 $out/bench_csr_8192.json
cmd="python bench.py --op csr --grid-n 8192 --steps 1 --warmup 3 --iters 10 --no-cpu-baseline --no-e2e"
$cmd > $out/plain.log 2>&1 &&
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:csr_ -s 30 -c 4 -f -o $out/csr $cmd > $out/ncu.log 2>&1
tail -3 $out/ncu.log
