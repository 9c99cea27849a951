out=gpurun_out
cap() {
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -c 1 \
      -o $out/cap_$1 -f python tools/profile_step.py 512 1 > $out/cap_$1.log 2>&1
  ncu -i $out/cap_$1.ncu-rep --page source --csv > $out/src_$1.csv 2>/dev/null
  rm -f $out/cap_$1.ncu-rep
}
cap attn_bwd7 'attn_bwd_kernel<.int.7, .int.256>'
cap dh_rmw   'gemm_tcgen05_kernel<.int.256, .bool.0, .int.3, .bool.1>'
ls -la $out/src_*.csv
