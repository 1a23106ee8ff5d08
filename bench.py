#!/usr/bin/env python
"""bench.py — scene-points/s of the AMContrast3D hot path (FPS + ball query/kNN + grouping + AM loss
fwd/bwd) on 1..8 B200, next to the reference algorithm on the host CPU cores.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's sm_100a kernels
    python bench.py --impl reference --steps K --warmup W     # CPU restatement of the reference (oracle/)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one data-parallel unit = BASELINE config 2: 8 S3DIS-shaped
scenes x 24 000 points through the PointNeXt-XL grouping operators and the AMContrast3D-AA loss,
forward + backward (amcontrast3d_b200/replay.py).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "scene-points/sec (FPS+kNN+group+AM loss fwd/bwd)"
UNIT = "scene-points/s"
WORKLOAD = "PointNeXt-XL grouping ops + AMContrast3D-AA loss fwd/bwd, unit = 8 x 24000-pt S3DIS-shaped scenes"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated in oracle/ (C for the operators, torch for the loss)
# ------------------------------------------------------------------------------------------
def cpu_path_replay(xyz, labels, k=16, seed=0):
    """One scene (1, N, 3) through the same call sequence as replay.PathReplay.step, on the host.
    Returns the loss value.  Executes oracle/ — allowed here only as the measured CPU baseline."""
    from amcontrast3d_b200.replay import XL, aa_args
    from oracle import loss_oracle as lo
    from oracle import ops_oracle as oo
    rng = np.random.default_rng(seed)
    B, N, _ = xyz.shape
    n, C = [N], [XL["width"]]
    for l in range(1, 5):
        n.append(n[-1] // XL["strides"][l])
        C.append(C[-1] * 2)
    F = [rng.standard_normal((B, C[l], n[l]), dtype=np.float32) for l in range(5)]
    p = [xyz]
    grouped = []
    for l in range(1, 5):
        idx, _ = oo.fps(p[l - 1], n[l])
        p.append(np.take_along_axis(p[l - 1], idx[:, :, None].astype(np.int64), axis=1))
        r = XL["radius"] * 2 ** (l - 1)
        bq = oo.ball_query(r, 32, p[l - 1], p[l])
        oo.group_points(np.ascontiguousarray(p[l - 1].transpose(0, 2, 1)), bq)
        grouped.append((oo.group_points(F[l - 1], bq), bq, n[l - 1]))
        for _ in range(XL["blocks"][l] - 1):
            bq = oo.ball_query(2 * r, 32, p[l], p[l])
            oo.group_points(np.ascontiguousarray(p[l].transpose(0, 2, 1)), bq)
            grouped.append((oo.group_points(F[l], bq), bq, n[l]))
    ups = []
    for l in range(4, 0, -1):
        d2, i3 = oo.three_nn(p[l - 1], p[l])
        recip = 1.0 / (np.sqrt(d2) + np.float32(1e-8))
        w = (recip / recip.sum(2, keepdims=True)).astype(np.float32)
        ups.append((oo.three_interpolate(F[l], i3, w), i3, w, n[l]))
    p_list = [pp.reshape(-1, 3) for pp in p[:4]]
    f_list = [torch.from_numpy(rng.standard_normal((B * n[s], C[s]), dtype=np.float32)).requires_grad_(True)
              for s in range(4)]
    sl = lo.make_stage_list([torch.from_numpy(np.ascontiguousarray(pp)) for pp in p_list], f_list)
    loss, _, _, _ = lo.contrast_head_forward(torch.from_numpy(labels.reshape(-1)), sl, 13, None, aa_args(k))
    loss.backward()
    for out, bq, nn in grouped:       # scatter-add backward of the 19 feature groupings
        oo.group_points_grad(out, bq, nn)
    for out, i3, w, m in ups:
        oo.three_interpolate_grad(out, i3, w, m)
    return float(loss.item())


def ref_gpu_path_replay(xyz_t, labels_t, k=16, seed=0, F=None, f_dec=None):
    """The same call sequence on the GPU with the REFERENCE's own CUDA kernels (oracle/_ref, compiled
    unmodified from the reference for sm_100) and the reference loss restated in torch on CUDA tensors —
    the "reference recompiled on the same B200" baseline of BASELINE.md §3.  xyz_t (B,N,3), labels_t (B,N)."""
    from amcontrast3d_b200.replay import XL, aa_args
    from oracle import loss_oracle as lo
    from oracle import ref_kernels as rk
    g = torch.Generator(device=xyz_t.device)
    g.manual_seed(seed)
    B, N, _ = xyz_t.shape
    n, C = [N], [XL["width"]]
    for l in range(1, 5):
        n.append(n[-1] // XL["strides"][l])
        C.append(C[-1] * 2)
    if F is None:
        F = [torch.randn((B, C[l], n[l]), device=xyz_t.device, generator=g) for l in range(5)]
    p = [xyz_t]
    grouped = []
    for l in range(1, 5):
        idx, _ = rk.fps(p[l - 1], n[l])
        p.append(torch.gather(p[l - 1], 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous())
        r = XL["radius"] * 2 ** (l - 1)
        bq = rk.ball_query(r, 32, p[l - 1], p[l])
        rk.group_points(p[l - 1].transpose(1, 2).contiguous(), bq)
        grouped.append((rk.group_points(F[l - 1], bq), bq, n[l - 1]))
        for _ in range(XL["blocks"][l] - 1):
            bq = rk.ball_query(2 * r, 32, p[l], p[l])
            rk.group_points(p[l].transpose(1, 2).contiguous(), bq)
            grouped.append((rk.group_points(F[l], bq), bq, n[l]))
    ups = []
    for l in range(4, 0, -1):
        d2, i3 = rk.three_nn(p[l - 1], p[l])
        recip = 1.0 / (torch.sqrt(d2) + 1e-8)
        w = (recip / recip.sum(2, keepdim=True)).contiguous()
        ups.append((rk.three_interpolate(F[l], i3, w), i3, w, n[l]))
    if f_dec is None:
        f_list = [torch.randn((B * n[s], C[s]), device=xyz_t.device, generator=g).requires_grad_(True) for s in range(4)]
    else:                                           # the SAME decoder features as the timed arm: the two losses must agree
        f_list = [f.detach().clone().requires_grad_(True) for f in f_dec]
    sl = lo.make_stage_list([pp.reshape(-1, 3).contiguous() for pp in p[:4]], f_list)
    loss, _, _, _ = lo.contrast_head_forward(labels_t.reshape(-1), sl, 13, None, aa_args(k), knn=rk.knnquery)
    loss.backward()
    for out, bq, nn in grouped:
        rk.group_points_grad(out, bq, nn)
    for out, i3, w, m in ups:
        rk.three_interpolate_grad(out, i3, w, m)
    torch.cuda.synchronize()
    return float(loss.item())


def time_ref_gpu(replay, k=16, our_loss=None):
    from oracle import ref_kernels as rk
    if not rk.available():
        return None
    try:
        xyz = replay.h_xyz.to(replay.device)                           # the batch every timed step ran on
        labels = replay.h_labels.to(replay.device)
        F = [f.detach() for f in replay.F]
        ref_gpu_path_replay(xyz, labels, k, F=F, f_dec=replay.f_dec)   # warm-up (allocator, module load)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = ref_gpu_path_replay(xyz, labels, k, F=F, f_dec=replay.f_dec)
        ms = 1e3 * (time.perf_counter() - t0)
    except Exception as e:                                             # a baseline must never break the bench
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    pts = replay.d_xyz.shape[0] * replay.d_xyz.shape[1]
    return {"value": pts / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": 1, "loss": loss,
            "loss_rel_diff": (abs(loss - our_loss) / abs(loss)) if our_loss is not None else None,
            "loss_note": "same xyz, labels and decoder features as the timed arm: loss_rel_diff is a parity check at the "
                         "headline config (bar 1e-5)",
            "kind": "the reference's own CUDA kernels (oracle/_ref: unmodified sources compiled for sm_100) for FPS / "
                    "ball_query / grouping / three_nn / interpolate / knnquery + the reference loss restated in torch, "
                    "same unit (8 x 24000 points) on the same GPU; host-timed, the reference launchers synchronise"}


def time_cpu_sample(steps, warmup, k=16, batch=8, points=24000, budget_s=150.0):
    """The reference algorithm on the host cores over the SAME unit as the GPU arm (`batch` scenes x `points`,
    flattened into one segment for the loss exactly as the encoder does).  Warm-up passes run on a small
    scene (they warm the thread pools, the allocator and the oracle library, not the caches of a 20 s pass);
    timed passes are full units, as many of the requested `steps` as fit in `budget_s` (at least one).
    -> (points/s, ms per pass, cores, passes run)"""
    from amcontrast3d_b200 import scenes
    from oracle import ops_oracle as oo
    cores = oo.host_threads()
    torch.set_num_threads(cores)
    wx, wl = scenes.batch_of_scenes(1, 4096, "surface")
    for _ in range(max(1, warmup)):
        cpu_path_replay(wx, wl, k)
    xyz, labels = scenes.batch_of_scenes(batch, points, "surface")
    times = []
    t_start = time.perf_counter()
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        cpu_path_replay(xyz, labels, k)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    ms = 1e3 * sum(times) / len(times)
    return batch * points / (ms / 1e3), ms, cores, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores, passes = time_cpu_sample(max(1, args.steps), args.warmup, args.k, args.batch, args.points)
    sample = (f"{passes} full pass(es) of the unit ({args.batch} scenes x {args.points} pts flattened into one segment, as on "
              "the GPU arm) through the full path: FPS, 19 ball queries, 38 groupings, three_nn/interpolate, AA loss "
              f"fwd+bwd, grouping/interpolate backward; {max(1, args.warmup)} warm-up pass(es) on a 4096-pt scene; passes "
              "capped by a 150 s budget")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": passes, "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch": args.batch, "points_per_scene": args.points, "k": args.k,
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# per-kernel accounting for the roofline object
# ------------------------------------------------------------------------------------------
def algorithmic_work(name, a):
    """(kind, amount) of algorithmic work of one C-ABI call from its leading int arguments.
    kind 'flop' -> FP32 FLOP (8 per pair: 3 sub, 3 mul, 2 add — BASELINE.md §4); 'byte' -> HBM bytes."""
    if name in ("amc3d_knnquery", "amc3d_knnquery_order"):
        n, m = a[0], a[1]
        return "flop", 8.0 * n * m
    if name in ("amc3d_ball_query", "amc3d_three_nn"):
        b, n, m = a[0], a[1], a[2]
        return "flop", 8.0 * b * n * m
    if name in ("amc3d_group_points_ws", "amc3d_group_points"):
        b, c, n, npnt, ns = a[:5]
        return "byte", 4.0 * b * (c * npnt * ns + c * n + npnt * ns)
    if name in ("amc3d_group_points_grad_ws", "amc3d_group_points_grad", "amc3d_group_points_grad_ws_set"):
        b, c, n, npnt, ns = a[:5]
        return "byte", 4.0 * b * (c * npnt * ns + 2 * c * n + npnt * ns)
    if name in ("amc3d_three_interpolate", "amc3d_three_interpolate_ws"):
        b, c, m, n = a[:4]
        return "byte", 4.0 * b * (c * n + c * m + 6 * n)
    if name in ("amc3d_three_interpolate_grad", "amc3d_three_interpolate_grad_ws", "amc3d_three_interpolate_grad_ws_set"):
        b, c, n, m = a[:4]
        return "byte", 4.0 * b * (c * n + 2 * c * m + 6 * n)
    return None, 0.0


def profile_step(replay):
    """One extra, untimed step with CUDA events around every C-ABI call -> per-entry totals."""
    from amcontrast3d_b200 import _capi
    torch.cuda.synchronize()
    _capi.PROFILE = []
    replay.step()
    torch.cuda.synchronize()
    prof, _capi.PROFILE = _capi.PROFILE, None
    agg = {}
    for name, e0, e1, a in prof:
        ms = e0.elapsed_time(e1)
        kind, amount = algorithmic_work(name, a)
        if name.startswith("amc3d_group_points") and a[1] < 8:
            name += "[xyz,C=3]"                   # direct kernel, not the TMA-staged one: listed separately
        d = agg.setdefault(name, {"calls": 0, "ms": 0.0, "flop": 0.0, "byte": 0.0})
        d["calls"] += 1
        d["ms"] += ms
        if kind:
            d[kind] += amount
    return agg


def fp32_peak_tflops():
    """Nominal non-tensor FP32 peak of a B200: 148 SMs x 128 lanes x 2 FLOP x max SM clock."""
    try:
        mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits",
                                    "-i", "0"], capture_output=True, text=True).stdout.split()[0])
    except Exception:
        mhz = 1965.0
    return 148 * 128 * 2 * mhz * 1e6 / 1e12


def measure_fp32_tflops(dev):
    """FFMA micro-benchmark (amc3d_fp32_probe: 8 independent FFMA chains per thread, 148 x 8 CTAs of 1024
    threads, no memory traffic), best of 5, CUDA events -> measured FP32-pipe TFLOP/s on this GPU."""
    import ctypes
    from amcontrast3d_b200 import _capi
    blocks, iters = 148 * 8, 16384
    out = torch.empty(blocks * 1024, dtype=torch.float32, device=dev)
    flop = ctypes.c_double(0.0)
    st = torch.cuda.current_stream(dev).cuda_stream
    best = 0.0
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _capi.call("amc3d_fp32_probe", iters, blocks, out.data_ptr(), ctypes.byref(flop), st)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = max(best, flop.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    _capi.LAUNCHES -= 7
    return best


def measure_hbm_directional(dev, gib=1.0):
    """Write-only, read-only and copy bandwidth of this GPU, live (torch fill / sum / copy_ over `gib` GiB, best of 5,
    CUDA events): context for the grouping pair, whose forward is ~97 % writes and whose backward is ~97 % reads of
    HBM — the contract's roofline denominator stays the copy figure of MEASURED_PEAKS.json."""
    n = int(gib * (1 << 30) / 4)
    a = torch.empty(n, dtype=torch.float32, device=dev)
    b = torch.empty(n, dtype=torch.float32, device=dev)
    res = {}
    for name, fn, nbytes in (("write_only", lambda: a.fill_(1.0), 4 * n), ("read_only", lambda: a.sum(), 4 * n),
                             ("copy", lambda: b.copy_(a), 8 * n)):
        best = 0.0
        for i in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        res[name + "_gbs"] = round(best, 1)
    del a, b
    torch.cuda.empty_cache()
    res["how"] = f"torch fill_ / sum / copy_ over {gib} GiB fp32, best of 5, CUDA events (library kernels, context only)"
    return res


def count_evaluated_pairs(replay):
    """One extra, untimed step with the counting instantiations of the culled search kernels
    (amc3d_search_stats): distance evaluations actually issued, per entry point."""
    from amcontrast3d_b200 import _capi
    counters = torch.zeros(8, dtype=torch.int64, device=replay.device)
    per_entry = {}
    last = [0]

    def after(name, a):
        if name not in ("amc3d_knnquery", "amc3d_knnquery_order", "amc3d_ball_query", "amc3d_three_nn"):
            return
        torch.cuda.synchronize()
        c = counters.cpu()
        tot = int(c[0] + c[2] + c[4])
        d = per_entry.setdefault(name, {"evaluated": 0, "brute_force_calls": 0})
        if tot == last[0]:                   # a brute-force call (small cloud): every pair is evaluated
            _, flop = algorithmic_work(name, a)
            d["evaluated"] += int(flop / 8)
            d["brute_force_calls"] += 1
        d["evaluated"] += tot - last[0]
        last[0] = tot

    torch.cuda.synchronize()
    _capi.call("amc3d_search_stats", counters.data_ptr())
    _capi.AFTER_CALL = after
    try:
        replay.step()
        torch.cuda.synchronize()
    finally:
        _capi.AFTER_CALL = None
        _capi.call("amc3d_search_stats", 0)
    _capi.LAUNCHES -= 2
    return per_entry


# ------------------------------------------------------------------------------------------
# the fused grouping -> conv -> BN -> ReLU -> max operator next to the module composition it replaces
# ------------------------------------------------------------------------------------------
def xl_layers(batch, points):
    """the 19 grouped convolutions of PointNeXt-XL at (batch, points): (kind, level, N, M, C_in, C_out, radius)"""
    from amcontrast3d_b200.replay import XL
    n, C = [points], [XL["width"]]
    for l in range(1, 5):
        n.append(n[-1] // XL["strides"][l])
        C.append(C[-1] * 2)
    layers = []
    for l in range(1, 5):
        r = XL["radius"] * 2 ** (l - 1)
        layers.append(("sa", l, n[l - 1], n[l], C[l - 1], C[l], r))
        layers += [("la", l, n[l], n[l], C[l], C[l], 2 * r)] * (XL["blocks"][l] - 1)
    return layers


def time_fused_operator(replay, dev, iters=3):
    """Forward + backward of the 19 grouped convolutions of one step (config 2), training-mode BatchNorm:
    (a) this package's fused operator (layers/fused.py, TF32 operands on tcgen05, no grouped tensor);
    (b) the composition the reference's modules execute — QueryAndGroup (this package's grouping kernels, i.e.
        already faster than the reference's), torch.cat, nn.Conv2d 1x1 through cuDNN (TF32 allowed, torch's
        default), nn.BatchNorm2d, ReLU, max — with torch autograd.  Ball-query indices are shared and not timed."""
    import torch.nn as nn
    from amcontrast3d_b200.layers import QueryAndGroup, ball_query
    from amcontrast3d_b200.layers.fused import fused_group_conv_bn_relu_max
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    p = replay._fps_chain(replay.d_xyz)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    work = []
    flop = 0.0
    for kind, l, N, M, cin, cout, r in xl_layers(replay.B, replay.N):
        sup, qry = (p[l - 1], p[l]) if kind == "sa" else (p[l], p[l])
        idx = ball_query(r, 32, sup, qry)
        conv = nn.Conv2d(cin + 3, cout, 1, bias=False).to(dev)
        bn = nn.BatchNorm2d(cout).to(dev)
        f = torch.randn((replay.B, cin, N), device=dev, generator=g).requires_grad_(True)
        go = torch.randn((replay.B, cout, M), device=dev, generator=g)
        work.append((qry, sup, idx, conv, bn, f, go, r))
        flop += 2.0 * replay.B * M * 32 * (cin + 3) * cout
    grouper = {}

    def fused_pass(backward):
        for qry, sup, idx, conv, bn, f, go, r in work:
            out = fused_group_conv_bn_relu_max(qry, sup, f, idx, conv.weight, bn, r, True, "tf32")
            if backward:
                out.backward(go)

    def composed_pass(backward):
        for qry, sup, idx, conv, bn, f, go, r in work:
            qg = grouper.setdefault(r, QueryAndGroup(r, 32, normalize_dp=True))
            dp, fj = qg(qry, sup, f, idx=idx)
            out = torch.relu(bn(conv(torch.cat([dp, fj], 1)))).max(-1)[0]
            if backward:
                out.backward(go)

    def timed(fn, backward):
        for _ in range(2):
            fn(backward)
        for w in work:
            w[5].grad = None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn(backward)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def timed_graph(fn, backward=True):
        """the same pass recorded into ONE CUDA graph and replayed: what it costs on the GPU once per-call Python
        and launch latency are out of the picture (a string with the reason if the pass cannot be captured)"""
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn(backward)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            for w in work:
                w[5].grad = None
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                fn(backward)
            for _ in range(2):
                g_.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                g_.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            del g_
            return round(ms, 3)
        except Exception as e:
            torch.cuda.synchronize()
            return f"not captured: {type(e).__name__}: {e}"[:160]

    res = {}
    try:
        f_fwd, f_all = timed(fused_pass, False), timed(fused_pass, True)
        c_fwd, c_all = timed(composed_pass, False), timed(composed_pass, True)
        f_graph = timed_graph(fused_pass)
        c_graph = timed_graph(composed_pass)
        f_fgraph = timed_graph(fused_pass, False)
        c_fgraph = timed_graph(composed_pass, False)
        gpu_fwd = f_fgraph if isinstance(f_fgraph, float) else f_fwd   # GPU time of the forward (eager if not captured)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        tf32_peak = float(peaks.get("bf16_tflops", 1660.0)) / 2.0       # the forward is ~2 ms of tensor work: burst figure
        res = {"layers": len(work), "conv_gflop_fwd": round(flop / 1e9, 1),
               "fused": {"fwd_ms": round(f_fwd, 3), "fwd_bwd_ms": round(f_all, 3), "fwd_graph_ms": f_fgraph,
                         "fwd_bwd_graph_ms": f_graph},
               "composition": {"fwd_ms": round(c_fwd, 3), "fwd_bwd_ms": round(c_all, 3), "fwd_graph_ms": c_fgraph,
                               "fwd_bwd_graph_ms": c_graph,
                               "what": "QueryAndGroup (this package's grouping kernels) + cat + cuDNN Conv2d 1x1 (TF32) + "
                                       "BatchNorm2d (training) + ReLU + max, torch autograd"},
               "speedup_fwd": round(c_fwd / f_fwd, 2), "speedup_fwd_bwd": round(c_all / f_all, 2),
               "speedup_fwd_bwd_graph": (round(c_graph / f_graph, 2) if isinstance(f_graph, float) and isinstance(c_graph, float)
                                         else None),
               "speedup_fwd_graph": (round(c_fgraph / f_fgraph, 2) if isinstance(f_fgraph, float) and isinstance(c_fgraph, float)
                                     else None),
               "fused_fwd_tflops": round(flop / (gpu_fwd * 1e-3) / 1e12, 1),
               "tensor_roofline": {"bound": "tensor", "achieved": round(flop / (gpu_fwd * 1e-3) / 1e12, 1), "peak": tf32_peak,
                                   "unit": "TFLOP/s", "frac": round(flop / (gpu_fwd * 1e-3) / 1e12 / tf32_peak, 4),
                                   "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst, cuBLAS) / 2: TF32 runs at half the bf16 rate",
                                   "peak_nominal": 1125.0, "frac_nominal": round(flop / (gpu_fwd * 1e-3) / 1e12 / 1125.0, 4),
                                   "note": "forward of the 19 layers as one CUDA graph = GPU time of the whole operator (transpose, "
                                           "weight tiles, the tcgen05 kernel, statistics, normalise); the tcgen05 kernel alone: "
                                           "profiles/r02_fused.md; the backward does 1/16 of these FLOPs by construction"},
               "grouped_tensor_bytes_avoided_fwd": int(sum(4.0 * replay.B * w[0].shape[1] * 32 * (w[5].shape[1] + 3) for w in work)),
               "mode": "eager (per-call Python included on both sides), TF32 operands on both sides, 3 iterations after 2 warm-ups"}
    except Exception as e:                                     # never break the headline line
        res = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    torch.backends.cudnn.allow_tf32 = prev
    del work
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    from amcontrast3d_b200 import _capi
    from amcontrast3d_b200 import dist as amdist
    from amcontrast3d_b200.replay import PathReplay
    import torch.distributed as tdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device for --impl ours: this package has no CPU path")
    rank, local, world = amdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _capi.load(build_if_missing=False)

    replay = PathReplay(batch=args.batch, n_points=args.points, device=dev, k=args.k, rank=rank,
                        prefetch=not args.no_prefetch)
    stats_layout = amdist.PackedStats(13)
    # N > 1: flat gradient buckets as DDP would reduce them; the step's packed statistics (loss terms, selected
    # points per stage, per-class tp / union / count, computed inside the step) ride in the tail of the last bucket
    buckets = (amdist.GradBuckets(int(args.grad_mb * 1e6 / 4), dev, tail_extra=stats_layout.size)
               if world > 1 and args.grad_mb > 0 else None)
    stats_sink = buckets.extra if buckets is not None else torch.zeros(stats_layout.size, device=dev)
    replay.stats_sink = stats_sink

    use_graph = not args.no_graph

    cur = {"replay": replay}

    def one_step(e2e=False, graph=False):
        replay = cur["replay"]
        nb = len(buckets.buckets) if buckets is not None else 0
        if world > 1 and buckets is not None and nb > 1:
            buckets.launch(0, nb - 1)            # ready during the backward in DDP: overlaps the step
        if graph:
            loss = replay.step_graph(replay.h_xyz, replay.h_labels) if e2e else replay.step_graph()
        elif e2e:
            xyz = replay.h_xyz.to(dev, non_blocking=True)
            labels = replay.h_labels.to(dev, non_blocking=True)
            loss = replay.step(xyz, labels)
        else:
            loss = replay.step()
        if world > 1:
            if buckets is not None:
                buckets.launch(max(nb - 1, 0), nb)   # the last (1 MB) bucket + the packed statistics: ready when the backward ends
                buckets.wait()
            else:
                amdist.all_reduce_packed(stats_sink)
        return loss

    def timed(n_steps, e2e, graph):
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(n_steps):
            loss = one_step(e2e, graph)
            if e2e:
                last = float(loss.item())        # device -> host read of the step's result
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            ms = t.item()
            tdist.barrier()
        return ms, last

    def timed_e2e(n_steps, graph):
        """End to end with the host in the loop the way a trainer runs it: every step copies its inputs from
        pinned host memory (async, stream-ordered before the step) and copies its loss back to pinned host
        memory; the host READS the loss of step i while step i+1 is already queued (one step of lag, as
        asynchronous logging does), so the device never idles waiting for Python."""
        host_loss = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        evs = [torch.cuda.Event() for _ in range(2)]
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(n_steps):
            loss = one_step(True, graph)
            host_loss[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
            evs[i & 1].record()
            if i > 0:
                evs[(i - 1) & 1].synchronize()
                last = float(host_loss[(i - 1) & 1][0])
        evs[(n_steps - 1) & 1].synchronize()
        last = float(host_loss[(n_steps - 1) & 1][0])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            ms = t.item()
            tdist.barrier()
        return ms, last

    warm = max(args.warmup, 3)
    for _ in range(warm):
        one_step(False, False)
    torch.cuda.synchronize()

    # eager arm (every operator call issued from Python each step) — reported beside the graph arm
    e2e_steps = max(1, min(args.steps, 10))
    # (Python's cyclic collector is paused for these loops, as timeit does: an eager step creates a few hundred
    # tensors and autograd nodes, and a generation-2 collection in the middle of ten steps is milliseconds)
    import gc
    gc.collect()
    gc.disable()
    try:
        eager_passes = [timed(e2e_steps, False, False)[0] / e2e_steps for _ in range(3)]
        one_step(True, False)
        e2e_passes = [timed(e2e_steps, True, False) for _ in range(2)]
    finally:
        gc.enable()
    eager_ms = min(eager_passes) * e2e_steps
    eager_e2e_ms, last_loss = min(e2e_passes, key=lambda t: t[0])
    eager = {"ms_per_step": eager_ms / e2e_steps, "e2e_ms_per_step": eager_e2e_ms / e2e_steps, "steps": e2e_steps,
             "passes_ms_per_step": [round(x, 3) for x in eager_passes],
             "e2e_passes_ms_per_step": [round(t[0] / e2e_steps, 3) for t in e2e_passes],
             "note": "best of the passes shown; Python's cyclic GC paused during the loops"}

    l0 = _capi.LAUNCHES
    one_step(False, False)
    launches = _capi.LAUNCHES - l0
    if use_graph:
        replay.capture(warmup=1)
        launches = replay.graph_launches
        for _ in range(warm):
            one_step(False, True)
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, _ = timed(args.steps, False, use_graph)
    one_step(True, use_graph)
    ms_e2e, last_loss = timed_e2e(args.steps, use_graph)
    ms_e2e_blocking, _ = timed(e2e_steps, True, use_graph)     # host blocks on loss.item() every step
    clocks = sampler.stop() if rank == 0 else None

    # the same unit WITHOUT the cross-step pipeline (every step starts with its own FPS chain), for comparison
    unpipelined = None
    if replay.prefetch:
        ru = PathReplay(batch=args.batch, n_points=args.points, device=dev, k=args.k, rank=rank, prefetch=False)
        ru.stats_sink = stats_sink
        cur["replay"] = ru
        for _ in range(warm):
            one_step(False, False)
        if use_graph:
            ru.capture(warmup=1)
            for _ in range(warm):
                one_step(False, True)
        ms_u, _ = timed(args.steps, False, use_graph)
        ms_ue, _ = timed_e2e(args.steps, use_graph)
        unpipelined = {"ms_per_step": ms_u / args.steps, "value": world * args.batch * args.points / (ms_u / args.steps / 1e3),
                       "e2e_ms_per_step": ms_ue / args.steps}
        cur["replay"] = replay
        del ru
        torch.cuda.empty_cache()

    # the same unit with every grouping of the encoder replaced by the fused grouping + conv + BN + ReLU + max operator
    # (compat tier 4): no grouped tensor exists in this step; it additionally executes the 19 convolutions
    fused_step = None
    if world == 1 and not args.no_fused:
        try:
            rf = PathReplay(batch=args.batch, n_points=args.points, device=dev, k=args.k, rank=rank, prefetch=not args.no_prefetch,
                            fused_conv=True)
            cur["replay"] = rf
            for _ in range(warm):
                one_step(False, False)
            if use_graph:
                rf.capture(warmup=1)
                for _ in range(warm):
                    one_step(False, True)
            ms_f, _ = timed(args.steps, False, use_graph)
            fused_step = {"ms_per_step": ms_f / args.steps, "value": args.batch * args.points / (ms_f / args.steps / 1e3),
                          "gpu_launches": int(getattr(rf, "graph_launches", 0)),
                          "what": "path replay with the 19 groupings replaced by the fused grouping -> 1x1 conv -> BatchNorm -> "
                                  "ReLU -> max operator (TF32), forward + backward: the 6.45 GB of grouped tensors are gone "
                                  "from the step, which now also contains the 19 convolutions (865 GFLOP forward) that the "
                                  "headline step leaves to the model"}
        except Exception as e:
            fused_step = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        cur["replay"] = replay
        try:
            del rf
        except NameError:
            pass
        torch.cuda.empty_cache()

    ms_step = ms_total / args.steps
    pts = args.batch * args.points
    value = world * pts / (ms_step / 1e3)
    e2e_value = world * pts / (ms_e2e / args.steps / 1e3)

    if rank != 0:
        if world > 1:
            tdist.destroy_process_group()
        return

    # kernel accounting (untimed extra step) and the CPU baseline (N = 1 only)
    # per-kernel accounting on a single-stream pass (no geometry streams, no pipeline): with the side streams
    # active, a call's event-to-event time also contains whatever ran next to it
    serial = PathReplay(batch=args.batch, n_points=args.points, device=dev, k=args.k, rank=rank,
                        geometry_stream=False, prefetch=False)
    for _ in range(2):
        serial.step()
    agg = profile_step(serial)
    evaluated = count_evaluated_pairs(serial)
    del serial
    torch.cuda.empty_cache()
    total_ms = sum(d["ms"] for d in agg.values())
    hbm_peak, peak_src = measured_peaks()
    fp32_nominal = fp32_peak_tflops()
    fp32_measured = measure_fp32_tflops(dev)
    hbm_dir = measure_hbm_directional(dev) if world == 1 else None
    kernels = []
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"entry": name, "calls": d["calls"], "ms": round(d["ms"], 4), "share": round(d["ms"] / total_ms, 4)}
        if d["flop"]:
            # FP32-pipe accounting of a search entry: what it EVALUATES (counted in the kernels) against the measured
            # FFMA peak; the all-pairs figure of BASELINE.md is kept only as the ratio of work the culling avoids
            ev = evaluated.get(name, {}).get("evaluated", 0)
            row["all_pairs"] = int(d["flop"] / 8)
            if ev:
                row["evaluated_pairs"] = ev
                row["evaluated_gpairs_per_s"] = round(ev / (d["ms"] * 1e-3) / 1e9, 2)
                row["evaluated_tflops"] = round(8.0 * ev / (d["ms"] * 1e-3) / 1e12, 3)
                row["frac_fp32_measured"] = round(row["evaluated_tflops"] / fp32_measured, 4) if fp32_measured else None
                row["pairs_avoided_factor"] = round(d["flop"] / 8 / ev, 2)
        if d["byte"]:
            row["gbs"] = round(d["byte"] / (d["ms"] * 1e-3) / 1e9, 1)
            row["frac_hbm_peak"] = round(row["gbs"] / hbm_peak, 4)
        kernels.append(row)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    # the dominant HBM-bound kernel: the larger of the two TMA-staged grouping kernels (19 launches each per step)
    KERNEL_OF = {"amc3d_group_points_ws": "group_fwd_tma_kernel", "amc3d_group_points_grad_ws": "group_bwd_tma_kernel",
                 "amc3d_group_points_grad_ws_set": "group_bwd_tma_kernel"}
    hbm_rows = [r for r in kernels if r["entry"] in KERNEL_OF and "gbs" in r]
    roofline = None
    if hbm_rows:
        dom = max(hbm_rows, key=lambda r: r["ms"])
        per_launch_ms = dom["ms"] / dom["calls"]
        t = (traffic or {}).get(dom["entry"])
        t_alg = ((traffic or {}).get("algorithmic") or {}).get(dom["entry"])
        roofline = {"kernel": f"{KERNEL_OF[dom['entry']]} via {dom['entry']}", "bound": "hbm", "achieved": dom["gbs"],
                    "peak": hbm_peak, "unit": "GB/s", "frac": dom["frac_hbm_peak"], "peak_source": peak_src,
                    "share_of_step": dom["share"], "launches_per_step": dom["calls"],
                    "ms_per_launch": round(per_launch_ms, 4),
                    "algorithmic_bytes_per_launch": round(agg[dom["entry"]]["byte"] / dom["calls"]),
                    "traffic": t,
                    "traffic_note": (f"ncu dram__bytes_read+write of one launch at the largest feature-grouping shape "
                                     f"{(traffic or {}).get('shape')} ({t_alg:.4g} algorithmic bytes there: ratio "
                                     f"{t / t_alg:.3f}); profiles/ncu_traffic.json, profiles/r02_kernels_ncu.md"
                                     if t and t_alg else None),
                    "note": "measured on a single-stream pass of the step; per-launch figures are means over the step's "
                            "launches of this entry point (transpose / "
                            "memset / accumulate launches of the call included in the time); FPS, the largest single "
                            "share, is latency-bound and reported in kernels[] as us per pick"}
    roofline_hbm = [{"kernel": r["entry"], "achieved": r["gbs"], "peak": hbm_peak, "unit": "GB/s",
                     "frac": r["frac_hbm_peak"], "share_of_step": r["share"],
                     "traffic": (traffic or {}).get(r["entry"])} for r in hbm_rows]
    fps_row = next((r for r in kernels if r["entry"] == "amc3d_furthest_point_sampling"), None)
    if fps_row is not None:
        picks = sum(replay.n[1:])
        fps_row["us_per_pick"] = round(1e3 * fps_row["ms"] / picks, 4)
        fps_row["picks"] = picks

    fused = time_fused_operator(replay, dev) if world == 1 and not args.no_fused else None
    ref_gpu = None
    if world == 1 and not args.no_cpu and not args.no_ref_gpu:
        ref_gpu = time_ref_gpu(replay, args.k, last_loss)
    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        v, ms_cpu, cores, passes = time_cpu_sample(1, 1, args.k, args.batch, args.points)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "ms": ms_cpu,
                        "sample": f"{passes} full pass of the same unit ({args.batch} scenes x {args.points} pts, one "
                                  "flattened segment), full path incl. loss fwd+bwd, after a warm-up pass on a 4096-pt scene"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch": args.batch, "points_per_scene": args.points, "k": args.k,
                       "units_per_rank_per_step": 1, "parallelism": f"dp{world} (scene units, no data-path collective)",
                       "l2": "working set per step (6.45 GB of grouped tensors) exceeds the 126 MB L2; no flush needed",
                       "grad_allreduce_mb": args.grad_mb if world > 1 else 0, "loss": last_loss,
                       "reduced_stats": {k: (v.tolist() if hasattr(v, "tolist") else v)
                                         for k, v in stats_layout.unpack(stats_sink.double().cpu()).items()}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": replay.h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "steps": args.steps,
                    "host_loop": "inputs copied from pinned host memory and the loss copied back every step; the host "
                                 "reads step i's loss while step i+1 is queued (one step of lag)",
                    "blocking_ms_per_step": ms_e2e_blocking / e2e_steps},
            "schedule": ("pipelined across steps: the geometry of batch i+1 (FPS chain, ball queries, three_nn, loss "
                         "labels/kNN/ambiguity) runs on side streams during the feature pass + backward of batch i "
                         "(every step still executes one full geometry pass and one full feature pass)"
                         if replay.prefetch else "every step starts with its own FPS chain"),
            "unpipelined": unpipelined,
            "gpu_launches": int(launches), "mode": "cuda-graph replay of the whole step" if use_graph else "eager",
            "eager": eager, "clocks": clocks, "roofline": roofline, "roofline_hbm": roofline_hbm,
            "fp32": {"fp32_tflops": round(fp32_measured, 2), "fp32_tflops_nominal": round(fp32_nominal, 2),
                     "how": "amc3d_fp32_probe: 8 independent FFMA chains/thread, 148x8 CTAs x 1024 threads, best of 5, CUDA "
                            "events; search rows report distance evaluations counted inside the kernels (8 FLOP each) "
                            "against it — the rest of their issue slots is box tests and top-k maintenance, see "
                            "profiles/r02_search_ncu.md for issue-slot utilisation"},
            "hbm_directional": hbm_dir,
            "kernels": kernels, "kernel_ms_per_step": round(total_ms, 3), "cpu_baseline": cpu_baseline,
            "ref_gpu": ref_gpu, "fused_operator": fused, "fused_step": fused_step}
    print(json.dumps(line), flush=True)
    if world > 1:
        tdist.destroy_process_group()


def main():
    # keep stdout for the ONE JSON line: libraries (NCCL's version banner, warnings) go to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--points", type=int, default=24000)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--grad-mb", type=float, default=166.3, help="flat gradient all-reduce per step (N>1): PointNeXt-XL FP32 grads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-CUDA-kernels-on-this-GPU baseline")
    ap.add_argument("--no-fused", action="store_true", help="skip the fused-operator vs module-composition leg")
    ap.add_argument("--no-prefetch", action="store_true", help="do not pipeline the FPS chain of the next batch into the current step")
    ap.add_argument("--no-graph", action="store_true", help="time the eager (per-call Python) step instead of the CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
