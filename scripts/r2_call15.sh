#!/bin/bash
# Round 2, GPU call 15 (two GPUs): the weak-scaling bench line on the final build - in-run sharded parity, value, and the
# end-to-end leg through the batch queue on a sharded plan (every rank queues its slab of b / x).
out=gpurun_out/r2_call15
mkdir -p $out
timeout -k 5 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 \
  bench.py --gpus 2 --steps 3 --warmup 3 > $out/weak_2gpu.json 2> $out/weak_2gpu.err; echo "weak rc=$?"
grep multigpu $out/weak_2gpu.err | tail -16
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_call15/weak_2gpu.json").read().strip().splitlines()[-1])
    print("value %.1f ms/step %.2f e2e %.1f serial %.1f batch_error %s parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"],
          d["e2e"]["one_call_per_step"]["value"], d["e2e"]["batch_error"], d.get("multi_gpu_parity")))
except Exception as e:
    print("parse error", e)
PY
tail -4 $out/weak_2gpu.err | cut -c1-300
