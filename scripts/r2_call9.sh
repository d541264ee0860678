#!/bin/bash
# Round 2, GPU call 9 (one GPU): experiment - one 15-warp CTA per SM on 840-column strips (B200CG_FUSED_CW=14).
out=gpurun_out/r2_call9
mkdir -p $out
B200CG_FUSED_CW=14 timeout -k 5 600 python -m pytest tests/test_single_sweep_gpu.py -m gpu -q --maxfail=5 2>&1 | tail -8 | tee $out/tests_cw14.log
: > $out/ab.txt
for rep in 1 2; do
  for v in "-" "B200CG_FUSED_CW=14"; do
    envs=""; [ "$v" != "-" ] && envs="$v"
    line=$(env $envs timeout -k 5 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extras 2>$out/err.txt | tail -1)
    python - "$v" "$line" >> $out/ab.txt <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2]); r = d["roofline"]
    print(f"{sys.argv[1]:24s} value {d['value']:.2f} even {r['update_kernel_even_iterations']['avg_launch_ms']:.4f} odd {r['avg_launch_ms']:.4f} mhz {d['clocks']['sm_mhz']}")
except Exception as exc:
    print(f"{sys.argv[1]:24s} FAILED {exc!r}")
PY
  done
done
cat $out/ab.txt; tail -3 $out/err.txt
