#!/bin/bash
# First multi-GPU call of round 2: the sharded single-sweep iteration (B200CG_SINGLE_SWEEP_SHARDED=1) - parity cases of
# tests/run_multigpu.py, then weak and strong scaling with it. Usage (gpurun --gpus N): scripts/round2_single_sweep_multigpu.sh N
n=${1:-2}
out=gpurun_out/r2_ss_mg
mkdir -p $out
B200CG_TEST_EXPERIMENTAL=1 timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
  --master-port 29561 tests/run_multigpu.py > $out/check_${n}ranks.log 2>&1
grep "multigpu\|MULTIGPU" $out/check_${n}ranks.log
for mode in weak strong; do
  B200CG_SINGLE_SWEEP_SHARDED=1 timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port 29562 bench.py --gpus $n --single-sweep 1 --scaling $mode --grid-n 16384 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline \
    > $out/${mode}_${n}gpu.json 2> $out/${mode}_${n}gpu.err
  tail -c 400 $out/${mode}_${n}gpu.json
done
