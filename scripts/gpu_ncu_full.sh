#!/bin/bash
# ncu --set full captures of single launches of the final kernels (raw metric pages -> profiles/r02_*_raw.csv)
mkdir -p gpurun_out
cap() {  # name, kernel regex, count, codec, precision, [batch]
  AC_PRECISION=$5 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -c $3 -o /tmp/$1 python scripts/profile_step.py 3 $4 $6 > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02_$1_raw.csv 2>/dev/null
  echo "$1: $(grep -c . gpurun_out/r02_$1_raw.csv) lines"
}
cap lstm lstm_tc 1 encodec exact
cap rvq rvq_encode_tc 1 encodec exact
cap dac_conv conv_tc 34 dac fp16 16
