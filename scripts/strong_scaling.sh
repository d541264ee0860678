#!/bin/bash
# Strong scaling of BASELINE.json configs[2]: the 16384^2 L-shaped grid sharded as row slabs over 1/2/4/8 GPUs of one box.
# Usage (on an 8-GPU box): scripts/strong_scaling.sh [outdir] [gpu counts, default "1 2 4 8"]
out=${1:-gpurun_out}
counts=${2:-"1 2 4 8"}
mkdir -p "$out"
common="--scaling strong --grid-n 16384 --iters 500 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
for g in $counts; do
  if [ "$g" = 1 ]; then
    timeout 300 python bench.py --gpus 1 $common > "$out/strong_1gpu.json" 2> "$out/strong_1gpu.err"
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29510 + g)) \
      bench.py --gpus $g $common > "$out/strong_${g}gpu.json" 2> "$out/strong_${g}gpu.err"
  fi
done
tail -c 300 "$out"/strong_*gpu.json
