"""N > 1 host logic on CPU: two gloo ranks shard whole units (flattened local batches), all-reduce the
packed statistics buffer and the flat gradient buckets (amcontrast3d_b200/dist.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from amcontrast3d_b200 import dist as amdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = amdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    units = amdist.shard_units(8, world, rank)
    layout = amdist.PackedStats(13)
    # every rank contributes loss = rank+1 and its unit ids as per-class counts
    tp = torch.zeros(13)
    for u in units:
        tp[u] += 1
    buf = layout.pack(rank + 1.0, 0.5, 0.25, [len(units)] * 4, tp, tp * 2, tp * 3)
    amdist.all_reduce_packed(buf)
    stats = layout.unpack(buf)
    # 1000 gradient floats in 250-float buckets, a 100-float tail bucket that also carries the packed statistics
    buckets = amdist.GradBuckets(1000, "cpu", bucket_mb=0.001, tail_mb=0.0004, tail_extra=layout.size)
    assert [b.numel() for b in buckets.buckets] == [250, 250, 250, 150, 100 + layout.size]
    assert buckets.extra.numel() == layout.size and buckets.extra.data_ptr() == buckets.flat[1000:].data_ptr()
    buckets.flat.fill_(rank + 1.0)
    buckets.launch()
    buckets.wait()
    out[rank] = (units, stats["loss_sum"], stats["n_selected"], stats["tp"].tolist(), float(buckets.flat[:1000].sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_shard_units_and_reduce():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    u0, loss0, nsel0, tp0, g0 = out[0]
    u1, loss1, nsel1, tp1, g1 = out[1]
    assert sorted(u0 + u1) == list(range(8)) and not set(u0) & set(u1)     # whole units, none shared
    assert loss0 == loss1 == 3.0                                           # 1 + 2
    assert nsel0 == nsel1 == [8, 8, 8, 8]
    assert tp0 == tp1 == [1] * 8 + [0] * 5                                 # every unit counted exactly once
    assert g0 == g1 == 3000.0                                              # (1 + 2) * 1000 elements


def test_single_process_is_identity():
    buf = amdist.PackedStats(3).pack(1.0, 2.0, 3.0, [1, 2, 3, 4], [1, 0, 0], [0, 1, 0], [0, 0, 1])
    before = buf.clone()
    assert amdist.all_reduce_packed(buf) is None
    assert torch.equal(buf, before)
    assert amdist.shard_units(5, 1, 0) == [0, 1, 2, 3, 4]
