# 2-GPU box: NCCL parity test, the whole GPU suite on GPU 0, bench at N=1 and N=2
set -x
nvidia-smi -L
python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_i_dist.log 2>&1; tail -5 gpurun_out/pytest_i_dist.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_i.log 2>&1; tail -3 gpurun_out/pytest_i.log
python bench.py --gpus 1 --no-other-configs > gpurun_out/bench_i_n1.json 2> gpurun_out/bench_i_n1.err; tail -c 300 gpurun_out/bench_i_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_i_n2.json 2> gpurun_out/bench_i_n2.err; tail -c 600 gpurun_out/bench_i_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_i_n2_ref.json 2>> gpurun_out/bench_i_n2.err
