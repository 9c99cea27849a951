#!/bin/bash
# ncu --set full captures of the step's dominant kernels, one launch each, summarised on the box (the .ncu-rep files are
# too large to travel back: only the raw-metric CSV lines are kept).   bash tools/ncu_capture.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
cap() {   # name, kernel regex (demangled, with template arguments)
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -c 1 \
      -o $out/cap_$1 -f python tools/profile_step.py 512 1 > $out/cap_$1.log 2>&1
  ncu -i $out/cap_$1.ncu-rep --page raw --csv > $out/ncu_${tag}_$1.csv 2>/dev/null
  rm -f $out/cap_$1.ncu-rep
}
cap inproj   'gemm_tcgen05_kernel<.int.256, .bool.0, .int.1, .bool.1>'
cap keyproj  'gemm_tcgen05_kernel<.int.256, .bool.0, .int.0, .bool.1>'
cap keyproj1 'gemm_tcgen05_kernel<.int.256, .bool.0, .int.2, .bool.1>'
cap dh_rmw   'gemm_tcgen05_kernel<.int.256, .bool.0, .int.3, .bool.1>'
cap attn_bwd7 'attn_bwd_kernel<.int.7, .int.256>'
cap attn_bwd1 'attn_bwd_kernel<.int.1, .int.256>'
cap pool_fwd7 'pool_fwd_kernel<.int.7, .int.256>'
ls -la $out/ncu_${tag}_*.csv
