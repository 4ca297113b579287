#!/bin/bash
# round 2 profiles: ncu launch lists (device time, DRAM bytes, tensor-pipe activity per launch) of one tuned full-size step
mkdir -p gpurun_out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for spec in "encodec exact" "encodec fp16" "dac fp16" "dac exact" "mimi fp16" "mimi exact"; do
  set -- $spec
  AC_PRECISION=$2 timeout 900 ncu --profile-from-start off --metrics $MET --clock-control none --csv --log-file gpurun_out/r02_${1}_${2}_launches.csv python scripts/profile_step.py 3 $1 > gpurun_out/r2m_$1_$2.log 2>&1
  echo "$1 $2: $(grep -c 'gpu__time_duration' gpurun_out/r02_${1}_${2}_launches.csv) metric rows"
done
ls -la gpurun_out | head -20
