#!/bin/bash
# Round 2, GPU call 19 (one GPU, < 1 min): ncu --set full of the default single-sweep kernels in the WIDE strip geometry
# (cg_fused_kernel<.., 14>, what 16384^2 runs since the geometry is chosen by slab size) - the committed capture of the
# x-touching flavour was taken in the 420-column geometry.
out=gpurun_out/r2_call19
mkdir -p $out
timeout -k 3 48 ncu --set full --clock-control none --import-source on -k regex:cg_fused_kernel -s 8 -c 2 -f -o $out/fused_wide \
  python scripts/ncu_target.py > $out/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $out/ncu.log; ls -la $out
