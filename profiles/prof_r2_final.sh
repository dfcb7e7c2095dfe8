#!/bin/bash
# round-2 evidence: getitem timing, ncu launch list of the bench command, ncu --set full of every kernel on the path
mkdir -p gpurun_out
python tools/time_getitem.py > gpurun_out/r2_getitem_timing.txt 2>&1; cat gpurun_out/r2_getitem_timing.txt
BENCH="python bench.py --gpus 1 --steps 4 --warmup 3 --reps 2 --no-cpu-baseline --no-other-configs --no-variants --no-e2e --no-torch-gpu-baseline --full-out gpurun_out/r2_prof_bench_full.json"
$BENCH > gpurun_out/r2_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_final_launches.csv $BENCH > gpurun_out/r2_prof_ncu_launches.log 2>&1
python tools/prof_detect.py 0.5 1 > gpurun_out/r2_prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_|graph_' -s 14 -c 7 -o gpurun_out/r2_final_nms python tools/prof_detect.py 0.5 1 > gpurun_out/r2_prof_ncu2.log 2>&1
python tools/prof_detect.py 0.001 80 > gpurun_out/r2_prof_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_|graph_' -s 14 -c 7 -o gpurun_out/r2_final_nms_nc80 python tools/prof_detect.py 0.001 80 > gpurun_out/r2_prof_ncu3.log 2>&1
python tools/prof_loss.py 1 > gpurun_out/r2_prof_plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'loss_|decode_' -s 18 -c 9 -o gpurun_out/r2_final_loss python tools/prof_loss.py 1 > gpurun_out/r2_prof_ncu4.log 2>&1
python tools/prof_loss.py 80 > gpurun_out/r2_prof_plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'loss_|decode_' -s 18 -c 9 -o gpurun_out/r2_final_loss_nc80 python tools/prof_loss.py 80 > gpurun_out/r2_prof_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_final_launches.csv; tail -n 2 gpurun_out/r2_prof_ncu*.log
