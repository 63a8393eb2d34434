"""Probe: achievable HBM bandwidth on this box for write-only / read-only / copy / 1:3 read:write mixes
(context for the roofline of the step kernel, whose traffic is ~25 % reads + 75 % writes)."""
import torch, time
dev = torch.device("cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
N = 1 << 28   # 256 Mi floats = 1 GiB
a = torch.empty(N, dtype=torch.float32, device=dev); b = torch.empty(N, dtype=torch.float32, device=dev)
t = timeit(lambda: a.fill_(1.0)); print("fill 1GiB      : %.1f GB/s (write only)" % (N*4/t/1e9))
t = timeit(lambda: a.zero_()); print("zero_ (memset) : %.1f GB/s (write only)" % (N*4/t/1e9))
t = timeit(lambda: b.copy_(a)); print("copy 1GiB      : %.1f GB/s (read+write bytes)" % (2*N*4/t/1e9))
t = timeit(lambda: a.sum()); print("sum 1GiB       : %.1f GB/s (read only)" % (N*4/t/1e9))
# 1:3 mix: read a quarter, write three quarters (out = expand of in)
src = torch.empty(N // 4, dtype=torch.float32, device=dev); dst = torch.empty((3, N // 4), dtype=torch.float32, device=dev)
t = timeit(lambda: torch.add(src.unsqueeze(0), 1.0, out=None) if False else dst.copy_(src.unsqueeze(0).expand(3, -1)))
print("1 read : 3 write: %.1f GB/s (read+write bytes)" % ((N + 3 * N) / 4 * 4 / t / 1e9 * 1.0))
# smaller working set comparable to the step kernel: write 214 MB
c = torch.empty(214 * 250000, dtype=torch.float32, device=dev)
t = timeit(lambda: c.fill_(2.0)); print("fill 214 MB    : %.1f GB/s" % (c.numel()*4/t/1e9))
