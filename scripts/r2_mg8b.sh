#!/bin/bash
# 8-GPU call, second part: the converged 16384^2 solve on 8 GPUs and the peer-exchange trace (strong and weak shapes).
n=${1:-8}
out=gpurun_out/r2_mg${n}
mkdir -p $out
tr() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "${@:2}"; }
tr 300 scripts/converged_runs.py --grid-n 16384 --modes single_sweep > $out/converged_16384.jsonl 2> $out/converged.err; echo "converged rc=$?"; cat $out/converged_16384.jsonl
tr 200 scripts/peer_trace.py --grid-n 16384 --iters 2000 > $out/peer_trace_strong.json 2> $out/peer_trace.err; echo "trace rc=$?"; head -c 900 $out/peer_trace_strong.json; echo
tr 200 scripts/peer_trace.py --grid-n 46342 --iters 1000 > $out/peer_trace_weak.json 2> $out/peer_trace_weak.err; echo "trace weak rc=$?"; head -c 900 $out/peer_trace_weak.json; echo
tail -n 3 $out/converged.err $out/peer_trace.err $out/peer_trace_weak.err
