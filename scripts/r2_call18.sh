#!/bin/bash
# Round 2, GPU call 18 (one GPU, < 1 min): BASELINE.json configs[1] (4096^2) and configs[0] (128^2) bench lines on the final
# build (their e2e legs now run through the right-hand-side queue).
out=gpurun_out/r2_call18
mkdir -p $out
timeout -k 5 60 python bench.py --grid-n 4096 --steps 10 --no-extras --no-cpu-baseline > $out/bench_4096.json 2> $out/bench_4096.err; head -c 150 $out/bench_4096.json; echo
timeout -k 5 40 python bench.py --grid-n 128 --iters 352 --steps 20 --no-extras --no-cpu-baseline > $out/bench_128.json 2> $out/bench_128.err; head -c 150 $out/bench_128.json; echo
python - <<'PY'
import json
for f in ("bench_4096", "bench_128"):
    try:
        d = json.loads(open(f"gpurun_out/r2_call18/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.2f e2e %.2f serial %.2f err %s" % (d["value"], d["e2e"]["value"], d["e2e"]["one_call_per_step"]["value"], d["e2e"]["batch_error"]))
    except Exception as e:
        print(f, "parse error", e)
PY
