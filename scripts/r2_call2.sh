#!/bin/bash
# Round 2, GPU call 2 (one GPU): stream-mix ceiling, parity of the restructured single-sweep kernel, stage-shape A/B, ncu.
out=gpurun_out/r2_call2
mkdir -p $out
timeout 120 scripts/stream_mix > $out/stream_mix.txt 2>&1; cat $out/stream_mix.txt
timeout -k 5 600 python -m pytest tests/test_single_sweep_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee $out/tests_ss.log
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "interrupt or dense or reupload or edge_cases or callback" 2>&1 | tail -8 | tee $out/tests_misc.log
: > $out/ab.txt
for rep in 1 2; do
  for v in "-" "B200CG_FUSED_NOX=1" "B200CG_FUSED_NOX=2" "B200CG_FUSED_X2=1" "B200CG_FUSED_X2=2"; do
    envs=""; [ "$v" != "-" ] && envs="$v"
    line=$(env $envs timeout -k 5 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-single-sweep-extra 2>$out/err.txt | tail -1)
    python - "$v" "$line" >> $out/ab.txt <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2]); r = d["roofline"]
    print(f"{sys.argv[1]:36s} value {d['value']:.2f} even {r['update_kernel_even_iterations']['avg_launch_ms']:.4f} odd {r['avg_launch_ms']:.4f} ss {r['single_sweep']} mhz {d['clocks']['sm_mhz']}")
except Exception as exc:
    print(f"{sys.argv[1]:36s} FAILED {exc!r}")
PY
  done
done
cat $out/ab.txt
cmd="python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu-baseline --no-e2e --no-single-sweep-extra"
$cmd > $out/plain.log 2>&1 &&
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:cg_fused_kernel -s 40 -c 2 -f -o $out/fused7 $cmd > $out/ncu.log 2>&1
tail -3 $out/ncu.log
