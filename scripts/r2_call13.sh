#!/bin/bash
# Round 2, GPU call 13 (one GPU, short): the new batch queue and the max-norm single sweep - their tests, the existing tests
# of the paths they touch, the default bench (e2e through the queue), the max-norm iteration's throughput.
out=gpurun_out/r2_call13
mkdir -p $out
timeout -k 5 330 python -m pytest tests/test_batch_gpu.py tests/test_single_sweep_gpu.py -m gpu -q --maxfail=30 --durations=12 \
  -k "batch or maxnorm or edge_cases or strips or wide_geometry" 2>&1 | tail -60 | tee $out/tests_new.log
timeout -k 5 150 python -m pytest tests/test_gpu_parity.py tests/test_dropin_gpu.py -m gpu -q --maxfail=30 --durations=8 \
  -k "maxnorm or exact_error or interrupt or dirichlet or msg or batch or callback or postprocess" 2>&1 | tail -40 | tee $out/tests_touched.log
timeout -k 5 120 python scripts/maxnorm_bench.py > $out/maxnorm_16384.jsonl 2> $out/maxnorm.err; cat $out/maxnorm_16384.jsonl | cut -c1-330; tail -3 $out/maxnorm.err
timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_call13/bench_default.json").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f (%s) serial %.1f batch_error %s" % (d["value"], d["e2e"]["value"], d["e2e"]["mode"][:24], d["e2e"]["one_call_per_step"]["value"], d["e2e"]["batch_error"]))
    print("roofline frac %.3f ms/step %.1f" % (d["roofline"]["frac"], d["ms_per_step"]))
except Exception as e:
    print("bench parse error", e)
PY
tail -5 $out/bench_default.err
