#!/bin/bash
# Round 2 multi-GPU call: sharded parity (in-run, bench.py), weak and strong scaling of both iteration schemes.
# Usage (gpurun --gpus N): scripts/r2_multigpu.sh N [extra bench args]
n=${1:-2}
out=gpurun_out/r2_mg${n}
mkdir -p $out
run() {  # name, args...
  name=$1; shift
  timeout -k 5 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $n "$@" > $out/$name.json 2> $out/$name.err
  echo "== $name rc=$?"; grep multigpu $out/$name.err | tail -12; tail -c 1500 $out/$name.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','scaling')}, 'e2e', d['e2e'] and d['e2e']['value'], 'parity', d.get('multi_gpu_parity'), d['config']['iteration'][:12], d['config']['parallelism'])
except Exception as e: print('no json', e)
"
}
run weak --steps 3 --warmup 3
run strong --scaling strong --steps 3 --warmup 3 --no-parity --no-e2e
run weak_two_sweep --single-sweep 2 --steps 3 --warmup 3 --no-parity --no-e2e
run strong_two_sweep --single-sweep 2 --scaling strong --steps 3 --warmup 3 --no-parity --no-e2e
tail -5 $out/*.err | tail -30
