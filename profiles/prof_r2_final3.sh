#!/bin/bash
# final code of the round: the two channels-last tests, then the ncu launch list of the bench command and ncu --set full of
# the NMS kernels (spatial / gather / edge / resolve changed since prof_r2_final.sh) and of the filter / score kernels
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_reference_suite.py -m gpu -q -x -k "channels_last" --timeout 1500 ) > gpurun_out/r2_cl_pytest.log 2>&1
tail -5 gpurun_out/r2_cl_pytest.log
BENCH="python bench.py --gpus 1 --steps 4 --warmup 3 --reps 2 --no-cpu-baseline --no-other-configs --no-variants --no-e2e --no-torch-gpu-baseline --full-out gpurun_out/r2_prof_bench_full.json"
$BENCH > gpurun_out/r2_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_final_launches.csv $BENCH > gpurun_out/r2_prof_ncu_launches.log 2>&1
python tools/prof_detect.py 0.5 1 > gpurun_out/r2_prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_|graph_' -s 12 -c 6 -o gpurun_out/r2_final_nms -f python tools/prof_detect.py 0.5 1 > gpurun_out/r2_prof_ncu2.log 2>&1
python tools/prof_detect.py 0.001 80 > gpurun_out/r2_prof_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_|graph_' -s 12 -c 6 -o gpurun_out/r2_final_nms_nc80 -f python tools/prof_detect.py 0.001 80 > gpurun_out/r2_prof_ncu3.log 2>&1
ls -la gpurun_out/r2_final_nms*.ncu-rep gpurun_out/r2_final_launches.csv; tail -n 2 gpurun_out/r2_prof_ncu*.log
