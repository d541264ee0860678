#!/bin/bash
# Round 2, GPU call 4 (one GPU): whole GPU suite, converged runs (plain CG vs multigrid-preconditioned), CSR line.
out=gpurun_out/r2_call4
mkdir -p $out
timeout -k 5 1800 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -40 | tee $out/tests.log
timeout -k 5 300 python scripts/converged_runs.py --grid-n 4096 > $out/converged_4096.jsonl 2> $out/converged_4096.err; cat $out/converged_4096.jsonl
timeout -k 5 400 python scripts/converged_runs.py --grid-n 16384 --modes multigrid,single_sweep > $out/converged_16384.jsonl 2> $out/converged_16384.err; cat $out/converged_16384.jsonl
timeout -k 5 300 python bench.py --op csr --grid-n 8192 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_csr_8192.json 2> $out/bench_csr.err; tail -c 1200 $out/bench_csr_8192.json
