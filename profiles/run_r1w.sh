# 2-GPU: CUDA-graph replay with the NCCL all-reduce captured inside the graph (experiment), bounded
set -x
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus 2 --graph-multi-gpu --no-variants > gpurun_out/bench_w_n2.json 2> gpurun_out/bench_w_n2.err; echo rc=$?; tail -c 500 gpurun_out/bench_w_n2.err
