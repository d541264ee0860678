#!/bin/bash
# Round 2, GPU call 16 (two GPUs): the sharded max-norm single sweep - bench.py's in-run parity cases (now 18, incl. msg /
# msg-2s / msg-w / stop-msg with callback records) and the weak-scaling value on this build.
out=gpurun_out/r2_call16
mkdir -p $out
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 \
  bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e > $out/weak_2gpu.json 2> $out/weak_2gpu.err; echo "weak rc=$?"
grep multigpu $out/weak_2gpu.err | tail -20
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_call16/weak_2gpu.json").read().strip().splitlines()[-1])
    print("value %.1f ms/step %.2f parity %s" % (d["value"], d["ms_per_step"], d.get("multi_gpu_parity")))
except Exception as e:
    print("parse error", e)
PY
grep -v multigpu $out/weak_2gpu.err | tail -5 | cut -c1-300
