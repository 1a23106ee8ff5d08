"""Which kernels share the GPU with the step when the gradient buckets are all-reduced next to it (N > 1)?
torch.profiler (CUPTI) on rank 0 over a few graph-replayed steps; prints NCCL kernel time per step, the time the
last bucket is exposed after the replay, and the step's own top kernels.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/prof_ddp.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from amcontrast3d_b200 import dist as amdist  # noqa: E402
from amcontrast3d_b200.replay import PathReplay  # noqa: E402

rank, local, world = amdist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
r = PathReplay(batch=8, n_points=24000, device=dev, k=16, rank=rank, prefetch=True)
layout = amdist.PackedStats(13)
buckets = amdist.GradBuckets(int(166.3e6 / 4), dev, tail_extra=layout.size)
r.stats_sink = buckets.extra
for _ in range(3):
    r.step()
r.capture(warmup=1)
nb = len(buckets.buckets)


def step():
    buckets.launch(0, nb - 1)
    r.step_graph()
    buckets.launch(nb - 1, nb)
    buckets.wait()


for _ in range(5):
    step()
torch.cuda.synchronize()
tdist.barrier()
steps = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type is not None and str(e.device_type).endswith("CUDA")]
    tot = {}
    for e in ev:
        d = tot.setdefault(e.name, [0, 0.0])
        d[0] += 1
        d[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    t0 = min(e.time_range.start for e in ev)
    t1 = max(e.time_range.end for e in ev)
    print(f"world {world}: {steps} steps in {(t1 - t0) / 1e3:.2f} ms of device timeline = {(t1 - t0) / 1e3 / steps:.3f} ms per step")
    nccl = {k: v for k, v in tot.items() if "nccl" in k.lower()}
    print("NCCL kernels:", {k[:60]: (v[0] // steps, round(v[1] / steps / 1e3, 3)) for k, v in nccl.items()}, "(launches, ms per step)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"   {v[1] / steps / 1e3:8.3f} ms/step  {v[0] // steps:4d}x  {k[:90]}")
tdist.barrier()
tdist.destroy_process_group()
