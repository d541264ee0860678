#!/bin/bash
# Round 2, the 8-GPU call: weak scaling line (with in-run parity and e2e), strong scaling of configs[2], the converged
# 16384^2 solve on 8 GPUs, the peer-exchange trace, and the e2e leg without NUMA binding for comparison.
n=${1:-8}
out=gpurun_out/r2_mg${n}
mkdir -p $out
tr() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "${@:2}"; }
tr 400 bench.py --gpus $n --steps 3 --warmup 3 > $out/weak.json 2> $out/weak.err; echo "weak rc=$?"; grep multigpu $out/weak.err | tail -12
tr 300 bench.py --gpus $n --scaling strong --steps 3 --warmup 3 --no-parity --no-e2e > $out/strong.json 2> $out/strong.err; echo "strong rc=$?"
tr 300 scripts/converged_runs.py --grid-n 16384 --modes single_sweep > $out/converged_16384.jsonl 2> $out/converged.err; echo "converged rc=$?"; cat $out/converged_16384.jsonl
tr 200 scripts/peer_trace.py --grid-n 16384 --iters 2000 > $out/peer_trace_strong.json 2> $out/peer_trace.err; echo "trace rc=$?"; head -c 700 $out/peer_trace_strong.json; echo
tr 300 bench.py --gpus $n --steps 2 --warmup 3 --no-parity --no-numa-bind > $out/weak_no_numa.json 2> $out/weak_no_numa.err; echo "no-numa rc=$?"
nvidia-smi topo -m > $out/topo.txt 2>&1
for f in weak strong weak_no_numa; do python - $out/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f"%d["value"], "e2e", d["e2e"] and round(d["e2e"]["value"],1), "parity", d.get("multi_gpu_parity") and d["multi_gpu_parity"].get("ok"), d["e2e"] and d["e2e"].get("host_numa_binding_rank0"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 $out/*.err | tail -40
