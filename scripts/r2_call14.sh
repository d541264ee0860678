#!/bin/bash
# Round 2, GPU call 14 (one GPU): the whole GPU suite + smoke + the default bench on the final build, then the launch list of
# the default bench and one ncu --set full capture of the max-norm single-sweep kernel (with u) at 16384^2.
out=gpurun_out/r2_call14
mkdir -p $out
timeout -k 5 420 python -m pytest tests -m gpu -q --maxfail=10 --durations=15 2>&1 | tail -45 | tee $out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/smoke.log
timeout -k 5 300 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"; tail -c 400 $out/bench_default.json; echo
timeout -k 5 120 python scripts/maxnorm_bench.py --grid-n 8192 > $out/maxnorm_8192.jsonl 2> $out/maxnorm_8192.err; cut -c1-200 $out/maxnorm_8192.jsonl
timeout -k 5 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > $out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:cg_fused_kernel -s 12 -c 2 -f -o $out/fused_maxn \
  python scripts/maxnorm_bench.py --iters 10 --reps 1 > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $out/ncu_full.log
ls -la $out
