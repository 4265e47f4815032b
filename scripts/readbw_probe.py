import torch, time
dev = torch.device("cuda:0")
b, n = 400_000, 4096
spec = torch.randn((b, 2, n), dtype=torch.float32, device=dev)   # [b][half][n floats] : half 0 = bins [0, N/2) as (re, im)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    z.record(); torch.cuda.synchronize()
    return a.elapsed_time(z) / reps
half = spec[:, 0, :]
t = timed(lambda: half.sum())
print("strided half read (16 KB of every 32 KB): %.3f ms  %.0f GB/s" % (t, b * n * 4 / t / 1e6))
full = spec.view(-1)[: b * n]
t = timed(lambda: full.sum())
print("contiguous read of the same bytes: %.3f ms  %.0f GB/s" % (t, b * n * 4 / t / 1e6))
t = timed(lambda: spec.sum())
print("contiguous read, 2x bytes: %.3f ms  %.0f GB/s" % (t, 2 * b * n * 4 / t / 1e6))
y = torch.empty_like(full)
t = timed(lambda: y.copy_(full))
print("copy (read+write): %.3f ms  %.0f GB/s" % (t, 2 * b * n * 4 / t / 1e6))
