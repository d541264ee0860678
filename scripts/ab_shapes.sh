#!/bin/bash
# A/B of launch shapes on one GPU: alternates the variants so that thermal drift hits both alike.
# Usage: scripts/ab_shapes.sh "ENV1=.. ENV2=.." "ENV1=.." ...   (each argument is one variant's environment; "-" = defaults)
out=gpurun_out/ab.txt
: > $out
for rep in 1 2; do
  for v in "$@"; do
    envs=""; [ "$v" != "-" ] && envs="$v"
    line=$(env $envs python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1)
    python - "$v" "$line" >> $out <<'PY'
import json, sys
d = json.loads(sys.argv[2]); r = d["roofline"]
print(f"{sys.argv[1]:40s} value {d['value']:.2f} dot {r['dot_kernel']['avg_launch_ms']:.4f} nox {r['update_kernel_even_iterations']['avg_launch_ms']:.4f} x2 {r['avg_launch_ms']:.4f} mhz {d['clocks']['sm_mhz']}")
PY
  done
done
cat $out
