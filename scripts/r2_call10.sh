#!/bin/bash
# Round 2, GPU call 10 (one GPU): CW = 7 vs 14 on mid-size grids.
out=gpurun_out/r2_call10
mkdir -p $out
: > $out/ab.txt
for n in 1024 4096 8192; do
  for v in "-" "B200CG_FUSED_CW=14"; do
    for rep in 1 2; do
    envs=""; [ "$v" != "-" ] && envs="$v"
    line=$(env $envs timeout -k 5 200 python bench.py --grid-n $n --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-extras 2>$out/err.txt | tail -1)
    python - "$n $v" "$line" >> $out/ab.txt <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    print(f"{sys.argv[1]:30s} value {d['value']:.2f} ms/step {d['ms_per_step']:.3f}")
except Exception as exc:
    print(f"{sys.argv[1]:30s} FAILED {exc!r}")
PY
    done
  done
done
cat $out/ab.txt
