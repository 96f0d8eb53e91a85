set -x
CMD="python bench.py --steps 2 --warmup 3 --batch 32768 --no-cpu-baseline --no-e2e --no-extra --min-timed-ms 0"
$CMD > gpurun_out/r2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
$CMD > gpurun_out/r2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 10 -c 9 -o gpurun_out/r2_conv_umma -f $CMD > gpurun_out/r2_ncu2.log 2>&1
CMD3="python bench.py --steps 2 --warmup 3 --batch 32768 --no-cpu-baseline --no-e2e --no-extra --min-timed-ms 0 --precision 3"
$CMD3 > gpurun_out/r2_plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_fast_launches.csv $CMD3 > gpurun_out/r2_ncu3.log 2>&1
CMD4="python tools/profile_u8.py"
$CMD4 > gpurun_out/r2_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 20 -c 1 -o gpurun_out/r2_stem_u8 -f $CMD4 > gpurun_out/r2_ncu4.log 2>&1
cat gpurun_out/r2_plain4.log | tail -3
