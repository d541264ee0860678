#!/bin/bash
# 8-GPU call on the final build (pipelined launches, ~50 ms graphs, wide strip geometry): weak line with in-run parity and
# e2e, strong line, the converged 16384^2 solve.
n=${1:-8}
out=gpurun_out/r2_mg${n}_final
mkdir -p $out
tr() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "${@:2}"; }
tr 400 bench.py --gpus $n --steps 3 --warmup 3 > $out/weak.json 2> $out/weak.err; echo "weak rc=$?"; grep multigpu $out/weak.err | tail -14
tr 300 bench.py --gpus $n --scaling strong --steps 3 --warmup 3 --no-parity --no-e2e > $out/strong.json 2> $out/strong.err; echo "strong rc=$?"
tr 300 scripts/converged_runs.py --grid-n 16384 --modes single_sweep > $out/converged_16384.jsonl 2> $out/converged.err; echo "converged rc=$?"; grep '^{' $out/converged_16384.jsonl | cut -c1-400
for f in weak strong; do python - $out/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f ms/step %.2f"%(d["value"],d["ms_per_step"]), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "parity", d.get("multi_gpu_parity") and (d["multi_gpu_parity"].get("ok"), d["multi_gpu_parity"].get("cases")))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -n 3 $out/weak.err $out/strong.err $out/converged.err | grep -v "^\*\*\*\|OMP_NUM" | tail -12
