# 2-GPU box, everything bounded by `timeout`: sparse/assign tests, NCCL parity, bench at N=2 (own + reference arm), bench N=1 quick
set -x
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -x -q -k "sparse or nccl or graph or build_targets or nchw" > gpurun_out/pytest_p.log 2>&1; tail -3 gpurun_out/pytest_p.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 > gpurun_out/bench_p_n2.json 2> gpurun_out/bench_p_n2.err; echo rc=$?; tail -c 400 gpurun_out/bench_p_n2.err
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_p_n2_ref.json 2>> gpurun_out/bench_p_n2.err; echo rc=$?
timeout 150 python bench.py --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_p_n1.json 2> gpurun_out/bench_p_n1.err; echo rc=$?
