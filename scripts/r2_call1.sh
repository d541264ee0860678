#!/bin/bash
# Round 2, GPU call 1 (one GPU): experimental single-sweep variants - parity, A/B on the headline workload, ncu capture.
out=gpurun_out/r2_call1
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
B200CG_TEST_EXPERIMENTAL=1 timeout -k 5 500 python -m pytest tests/test_single_sweep_gpu.py -m gpu -q 2>&1 | tail -15 | tee $out/tests.log
: > $out/ab.txt
for rep in 1 2; do
  for v in "-" "B200CG_FUSED_DELTA=1" "B200CG_SHAPE_FUSED=1" "B200CG_SHAPE_FUSED=2" "B200CG_FUSED_DELTA=1 B200CG_SHAPE_FUSED=2" "B200CG_FUSED_DELTA=1 B200CG_SHAPE_FUSED=1"; do
    envs=""; [ "$v" != "-" ] && envs="$v"
    line=$(env $envs timeout -k 5 120 python bench.py --single-sweep 1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>$out/err.txt | tail -1)
    python - "$v" "$line" >> $out/ab.txt <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2]); r = d["roofline"]
    print(f"{sys.argv[1]:48s} value {d['value']:.2f} even {r['update_kernel_even_iterations']['avg_launch_ms']:.4f} odd {r['avg_launch_ms']:.4f} mhz {d['clocks']['sm_mhz']}")
except Exception as exc:
    print(f"{sys.argv[1]:48s} FAILED {exc!r}")
PY
  done
done
cat $out/ab.txt
cmd="python bench.py --single-sweep 1 --steps 1 --warmup 3 --iters 20 --no-cpu-baseline --no-e2e"
$cmd > $out/plain.log 2>&1 &&
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:cg_fused_kernel -s 40 -c 2 -f -o $out/fused $cmd > $out/ncu.log 2>&1
tail -3 $out/ncu.log
