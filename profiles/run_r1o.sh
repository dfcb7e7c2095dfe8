# scaling: bench at N = 2, 4, 8 GPUs of one box (NCCL), plus the NCCL parity test
set -x
nvidia-smi -L | head -8
python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_o_dist.log 2>&1; tail -3 gpurun_out/pytest_o_dist.log
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N > gpurun_out/bench_o_n$N.json 2> gpurun_out/bench_o_n$N.err; tail -c 300 gpurun_out/bench_o_n$N.err
done
python bench.py --gpus 1 --no-other-configs --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_o_n1.json 2> gpurun_out/bench_o_n1.err
