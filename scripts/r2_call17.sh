#!/bin/bash
# Round 2, GPU call 17 (one GPU, ~1 min): sanity of the last build on one GPU - smoke, the single-sweep / max-norm / batch /
# solve-parity tests.
out=gpurun_out/r2_call17
mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/smoke.log
timeout -k 5 100 python -m pytest tests/test_single_sweep_gpu.py tests/test_batch_gpu.py tests/test_gpu_parity.py -m gpu -q --maxfail=10 \
  -k "maxnorm or reference_fixtures or batch or interrupt or solve_parity or exact_error or fixed_iteration or edge_cases" 2>&1 | tail -8 | tee $out/tests.log
