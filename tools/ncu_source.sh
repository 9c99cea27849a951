#!/bin/bash
# SASS-level stall samples of one kernel of the train step: bash tools/ncu_source.sh <name> '<demangled kernel regex>'
out=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -c 1 \
    -o $out/cap_$1 -f python tools/profile_step.py 512 1 > $out/cap_$1.log 2>&1
ncu -i $out/cap_$1.ncu-rep --page source --csv > $out/src_$1.csv 2>/dev/null
ncu -i $out/cap_$1.ncu-rep --page raw --csv > $out/raw_$1.csv 2>/dev/null
rm -f $out/cap_$1.ncu-rep
ls -la $out/src_$1.csv $out/raw_$1.csv
