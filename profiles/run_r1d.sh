set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_d.log 2>&1; tail -3 gpurun_out/pytest_d.log
python bench.py --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -c 300 gpurun_out/bench_d.err
python - <<'PY' > gpurun_out/l2gran.log 2>&1
import torch, ctypes
rt = ctypes.CDLL("libcudart.so.12")
v = ctypes.c_size_t(0)
torch.zeros(1, device="cuda")
print("get", rt.cudaDeviceGetLimit(ctypes.byref(v), 5), v.value)   # cudaLimitMaxL2FetchGranularity = 0x05
import time
def bw(fn, nbytes, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return nbytes / (a.elapsed_time(b) / n * 1e-3) / 1e9
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); y = torch.empty_like(x)
print("memset GB/s", bw(lambda: x.zero_(), x.numel() * 4))
print("copy GB/s (r+w)", bw(lambda: y.copy_(x), x.numel() * 8))
print("read-only sum GB/s", bw(lambda: x.sum(), x.numel() * 4))
# strided 4B reads at 340B stride
idx = None
s = x[: (1 << 28) // 85 * 85].view(-1, 85)[:, 4]
print("strided obj column read: rows/s G", bw(lambda: s.sum(), s.numel() * 32) , "(GB/s if 32 B per row)")
for g in (32, 64, 128):
    print("set", g, rt.cudaDeviceSetLimit(5, ctypes.c_size_t(g)), end=" ")
    print("get", rt.cudaDeviceGetLimit(ctypes.byref(v), 5), v.value, end=" ")
    print("strided GB/s@32B/row", bw(lambda: s.sum(), s.numel() * 32))
PY
cat gpurun_out/l2gran.log
