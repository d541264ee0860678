#!/bin/bash
# First 1-GPU call of round 2 for the single-sweep kernel: parity of the variants that have only seen the CPU model,
# then an A/B of them on the headline workload. Usage (gpurun, one GPU): scripts/round2_single_sweep_1gpu.sh
out=gpurun_out/r2_ss
mkdir -p $out
B200CG_TEST_EXPERIMENTAL=1 timeout -k 5 400 python -m pytest tests/test_single_sweep_gpu.py -m gpu -q 2>&1 | tail -15 | tee $out/tests.log
: > $out/ab.txt
for rep in 1 2; do
  for v in "-" "B200CG_FUSED_DELTA=1" "B200CG_SHAPE_FUSED=1" "B200CG_SHAPE_FUSED=2" "B200CG_FUSED_DELTA=1 B200CG_SHAPE_FUSED=2"; do
    envs=""; [ "$v" != "-" ] && envs="$v"
    line=$(env $envs timeout -k 5 120 python bench.py --single-sweep 1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1)
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
