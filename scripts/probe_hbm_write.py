"""HBM bandwidth of a pure write stream, a pure read stream and a copy on this GPU (torch fill_ / sum / copy_ over 8 GiB),
for the write-heavy training kernels' roofline."""
import torch

n = 2 * 1024 ** 3          # 2 Gi floats = 8 GiB
x = torch.empty(n, device="cuda")
y = torch.empty(n, device="cuda")


def timeit(fn, nbytes, name, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:24s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s")


timeit(lambda: x.fill_(1.0), 4 * n, "write only (fill_)")
timeit(lambda: x.sum(), 4 * n, "read only (sum)")
timeit(lambda: y.copy_(x), 8 * n, "copy (read + write)")
