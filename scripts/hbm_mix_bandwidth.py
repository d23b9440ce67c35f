"""HBM bandwidth of one B200 for different read:write mixes (torch kernels, CUDA events, 2 GiB tensors): the copy kernel of
this library moves 11.5 GB of reads and 25 GB of writes per generation, so the 1:1 copy peak is not its exact ceiling."""
import torch
n = 1 << 29  # 2 GiB of float32
a, b, c = (torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(3))
a.normal_(); b.normal_()


def timeit(f, bytes_moved, name):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        f()
    e.record(); e.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"{name}: {bytes_moved / ms / 1e6:.0f} GB/s")


timeit(lambda: c.copy_(a), 2 * n * 4, "copy 1 read : 1 write")
timeit(lambda: c.zero_(), n * 4, "write only (zero_)")
timeit(lambda: a.sum(), n * 4, "read only (sum)")
timeit(lambda: torch.add(a, b, out=c), 3 * n * 4, "2 reads : 1 write (add)")
