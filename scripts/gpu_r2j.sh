#!/bin/bash
mkdir -p gpurun_out
M='gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|sm__pipe_tensor_subpipe_hmma_cycles_active|sm__inst_executed_pipe|smsp__issue_active.avg.pct|smsp__inst_executed.sum|sm__warps_active.avg.pct|l1tex__data_pipe_lsu_wavefronts|smsp__average_warp|smsp__average_warps_issue_stalled|launch__registers|sm__throughput|gpu__dram_throughput|l1tex__throughput|lts__throughput|smsp__inst_executed_pipe_xu|sm__inst_executed_pipe_xu|smsp__warp_issue_stalled.*_per_warp_active'
AC_PRECISION=fp16 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:resunit_tc -c 4 -o /tmp/enc python scripts/profile_step.py 3 encodec > gpurun_out/r2j_ncu1.log 2>&1; tail -1 gpurun_out/r2j_ncu1.log
ncu -i /tmp/enc.ncu-rep --page raw --csv > gpurun_out/r02_encodec_fp16_resunit_raw.csv
ncu -i /tmp/enc.ncu-rep --page source --csv > gpurun_out/r02_encodec_fp16_resunit_source.csv 2>/dev/null
AC_PRECISION=fp16 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:resunit_tc -c 14 -o /tmp/dac python scripts/profile_step.py 3 dac 16 > gpurun_out/r2j_ncu2.log 2>&1; tail -1 gpurun_out/r2j_ncu2.log
ncu -i /tmp/dac.ncu-rep --page raw --csv > gpurun_out/r02_dac_fp16_resunit_raw.csv
ls -la gpurun_out/
