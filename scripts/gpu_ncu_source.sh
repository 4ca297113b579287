#!/bin/bash
# source-level ncu captures (SASS hot spots) of the fused unit and the tap-GEMM kernel after the instruction diet
mkdir -p gpurun_out
AC_PRECISION=fp16 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:resunit_tc -c 2 -o /tmp/enc_ru python scripts/profile_step.py 3 encodec > gpurun_out/r2p_ncu1.log 2>&1; tail -1 gpurun_out/r2p_ncu1.log
ncu -i /tmp/enc_ru.ncu-rep --page raw --csv > gpurun_out/r2p_encodec_fp16_resunit_raw.csv
ncu -i /tmp/enc_ru.ncu-rep --page source --csv > /tmp/src1.csv 2>/dev/null; python scripts/ncu_source_summary.py /tmp/src1.csv gpurun_out/r2p_encodec_fp16_resunit_source.csv
AC_PRECISION=fp16 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -c 3 -o /tmp/enc_ct python scripts/profile_step.py 3 encodec > gpurun_out/r2p_ncu2.log 2>&1; tail -1 gpurun_out/r2p_ncu2.log
ncu -i /tmp/enc_ct.ncu-rep --page raw --csv > gpurun_out/r2p_encodec_fp16_conv_raw.csv
ncu -i /tmp/enc_ct.ncu-rep --page source --csv > /tmp/src2.csv 2>/dev/null; python scripts/ncu_source_summary.py /tmp/src2.csv gpurun_out/r2p_encodec_fp16_conv_source.csv
AC_PRECISION=fp16 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:resunit_tc -c 14 -o /tmp/dac_ru python scripts/profile_step.py 3 dac 16 > gpurun_out/r2p_ncu3.log 2>&1; tail -1 gpurun_out/r2p_ncu3.log
ncu -i /tmp/dac_ru.ncu-rep --page raw --csv > gpurun_out/r2p_dac_fp16_resunit_raw.csv
ncu -i /tmp/dac_ru.ncu-rep --page source --csv > /tmp/src3.csv 2>/dev/null; python scripts/ncu_source_summary.py /tmp/src3.csv gpurun_out/r2p_dac_fp16_resunit_source.csv
ls -la gpurun_out/r2p*
