"""What plain streaming kernels reach on this GPU: write-only (fill), read-only (sum), copy — the ceilings the
grouping gather (write stream) and scatter-add (read stream) can be held against."""
import torch

n = 786_432_000 // 4
x = torch.empty(n, device="cuda")
y = torch.empty(n, device="cuda")


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


gb = n * 4 / 1e9
print(f"fill  (write {gb:.2f} GB): {gb / t(lambda: x.fill_(1.0)) * 1e3:.0f} GB/s")
print(f"memset(write {gb:.2f} GB): {gb / t(lambda: x.zero_()) * 1e3:.0f} GB/s")
print(f"sum   (read  {gb:.2f} GB): {gb / t(lambda: x.sum()) * 1e3:.0f} GB/s")
print(f"copy  (r+w {2 * gb:.2f} GB): {2 * gb / t(lambda: y.copy_(x)) * 1e3:.0f} GB/s")
