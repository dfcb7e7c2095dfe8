# 2-GPU: final bench code (graph replay with forked loss branch, NCCL all-reduce captured on the branch stream), bounded
set -x
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --no-variants > gpurun_out/bench_w2_n2.json 2> gpurun_out/bench_w2_n2.err; echo rc=$?; tail -c 400 gpurun_out/bench_w2_n2.err
