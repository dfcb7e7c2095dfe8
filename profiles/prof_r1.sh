set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'nms_mask_kernel|loss_main_kernel|nms_scan_kernel|nms_sort_kernel|filter_emit_kernel' -s 15 -c 5 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/ncu2.log
