"""Data-parallel harness for the hot path: one process per GPU, NCCL over NVLink for plumbing only.

The path shards by scene with no exchange step inside any kernel.  Because the encoder flattens
the local batch into ONE kNN segment (offset = [B*n], pointnext_AA.py:461), the global
max(count) of the ambiguity function (AEF/ambiguity.py:14) and the per-stage mean over selected
points (MarginContrast.py:257) are all per *local batch*; the unit of work is therefore one
flattened local batch ("unit", 8 scenes in BASELINE configs 2 and 4), and results are invariant
to the number of ranks only if units stay intact.

Collectives per step (the reference: DDP gradient all-reduce main_AA.py:151, metric all-reduces
:461,496,502,507):
  * ONE packed all-reduce of [loss_sum, ce, am, n_selected[4], tp[ncls], union[ncls], count[ncls]]
    instead of the reference's separate small all-reduces;
  * the gradient all-reduce as flat buckets on a side stream, overlapped with the backward.
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).
    Returns (rank, local_rank, world_size); no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def shard_units(num_units: int, world_size: int, rank: int) -> List[int]:
    """DistributedSampler-style round-robin of whole units over ranks (dataset/build.py:79)."""
    return list(range(rank, num_units, world_size))


class PackedStats:
    """Layout of the single per-step all-reduce buffer (float64 so integer counts stay exact)."""

    def __init__(self, num_classes: int, stages: int = 4):
        self.ncls, self.stages = num_classes, stages
        self.size = 3 + stages + 3 * num_classes

    def pack(self, loss_sum, ce, am, n_selected, tp, union, count, device=None):
        buf = torch.zeros(self.size, dtype=torch.float64, device=device)
        buf[0], buf[1], buf[2] = float(loss_sum), float(ce), float(am)
        buf[3:3 + self.stages] = torch.as_tensor(n_selected, dtype=torch.float64)
        o = 3 + self.stages
        for part in (tp, union, count):
            buf[o:o + self.ncls] = torch.as_tensor(part, dtype=torch.float64)
            o += self.ncls
        return buf

    def pack_device(self, loss_sum, ce, am, n_selected_dev, tp, union, count):
        """Same, from device tensors, without a host sync."""
        parts = [torch.stack([loss_sum.detach().double().reshape(()), ce.detach().double().reshape(()),
                              am.detach().double().reshape(())]),
                 n_selected_dev.double().reshape(-1), tp.double().reshape(-1), union.double().reshape(-1),
                 count.double().reshape(-1)]
        return torch.cat(parts)

    def unpack(self, buf):
        s, n = self.stages, self.ncls
        o = 3 + s
        return dict(loss_sum=buf[0].item(), ce=buf[1].item(), am=buf[2].item(),
                    n_selected=buf[3:3 + s].round().long().tolist(),
                    tp=buf[o:o + n].round().long(), union=buf[o + n:o + 2 * n].round().long(),
                    count=buf[o + 2 * n:o + 3 * n].round().long())


def all_reduce_packed(buf: torch.Tensor, async_op: bool = False):
    """SUM-all-reduce of the packed buffer (identity for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=async_op)
    return None


class GradBuckets:
    """Flat gradient buckets all-reduced on a side stream so the transfer overlaps the backward
    (what DDP does for the 41.58 M-parameter PointNeXt-XL: 166.3 MB FP32).
    Buckets are sized for launch latency (NVSwitch makes bandwidth uniform), default 32 MB."""

    def __init__(self, numel: int, device, bucket_mb: float = 32.0, tail_mb: float = 1.0, tail_extra: int = 0):
        """`tail_mb`: size of the LAST bucket — the one that only becomes ready when the backward ends and whose
        all-reduce is therefore exposed; DDP makes the corresponding (first-allocated) bucket 1 MB for the same
        reason.  `tail_extra` floats are appended to it (`self.extra`): the step's packed statistics ride along
        with the last gradient bucket instead of paying for a collective of their own."""
        self.flat = torch.zeros(numel + tail_extra, dtype=torch.float32, device=device)
        per = max(1, int(bucket_mb * 1e6 / 4))
        tail = min(numel, max(1, int(tail_mb * 1e6 / 4)))
        body = numel - tail
        self.buckets = [self.flat[i:min(i + per, body)] for i in range(0, body, per)]
        self.buckets.append(self.flat[body:])
        self.extra = self.flat[numel:] if tail_extra else None
        self.stream = torch.cuda.Stream(device=device) if torch.device(device).type == "cuda" else None
        self._work = []

    def launch(self, first: int = 0, last: int | None = None):
        """Issue the all-reduces of buckets [first, last) (averaging by world size afterwards is left to
        the optimizer step).  DDP's buckets become ready one after another during the backward, so all but
        the last overlap compute; a caller that replays the step as one CUDA graph models that by
        launching buckets [0, n-1) before the replay and bucket n-1 after it."""
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            return
        todo = self.buckets[first:last]
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for b in todo:
                    self._work.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=True))
        else:
            for b in todo:
                self._work.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=True))

    def wait(self):
        for w in self._work:
            w.wait()
        self._work = []
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
