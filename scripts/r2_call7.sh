#!/bin/bash
# Round 2, GPU call 7 (one GPU): the branch-free FULL path of the single-sweep kernel - parity, bench, ncu.
out=gpurun_out/r2_call7
mkdir -p $out
timeout -k 5 900 python -m pytest tests/test_single_sweep_gpu.py tests/test_gpu_parity.py -m gpu -q --maxfail=5 -k "not config4 and not full_size" 2>&1 | tail -30 | tee $out/tests.log
for rep in 1 2; do
timeout -k 5 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $out/bench_$rep.json 2> $out/bench_$rep.err; python - $out/bench_$rep.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print("value %.2f e2e %.2f even %.4f odd %.4f mhz %s"%(d["value"], d["e2e"]["value"], r["update_kernel_even_iterations"]["avg_launch_ms"], r["avg_launch_ms"], d["clocks"]["sm_mhz"]))
PY
done
cmd="python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu-baseline --no-e2e --no-extras"
$cmd > $out/plain.log 2>&1 &&
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:cg_fused_kernel -s 40 -c 2 -f -o $out/fused_bf $cmd > $out/ncu.log 2>&1
tail -2 $out/ncu.log
