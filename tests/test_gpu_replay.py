"""Path replay (amcontrast3d_b200/replay.py) on the GPU: the CUDA-graph replay of a step must reproduce the
eager step, the AMContrast3D++ (MM) replay must run with refinement + ignore_index, and the full-size
(BASELINE config 2) calls are checked through size-independent properties."""
import numpy as np
import pytest
import torch

from amcontrast3d_b200 import scenes
from oracle import ops_oracle as oo

pytestmark = pytest.mark.gpu


def _grads(r):
    return [t.grad.detach().clone() for t in r.F + r.f_dec if t.grad is not None]


def test_graph_replay_matches_eager_step():
    from amcontrast3d_b200.replay import PathReplay
    r = PathReplay(batch=2, n_points=4096, k=16)
    r.step()                                           # sizes the synthetic upstream-gradient buffer
    loss_e = r.step().detach().clone()
    g_e = _grads(r)
    r.capture(warmup=2)
    assert r.graph_launches > 100                      # the whole step is in the graph
    for _ in range(2):
        loss_g = r.step_graph()
    torch.cuda.synchronize()
    assert torch.isfinite(loss_g)
    assert abs(loss_g.item() - loss_e.item()) <= 1e-6 * abs(loss_e.item())
    g_g = _grads(r)
    assert len(g_g) == len(g_e) and len(g_g) >= 8
    for a, b in zip(g_g, g_e):
        assert (a - b).norm() <= 1e-5 * b.norm()          # scatter-add atomics reorder
    # new host inputs flow through the static buffers
    xyz2, lab2 = scenes.batch_of_scenes(2, 4096, "surface", first_scene=11)
    h_xyz, h_lab = torch.from_numpy(xyz2).pin_memory(), torch.from_numpy(lab2).pin_memory()
    loss2 = r.step_graph(h_xyz, h_lab).item()
    r2 = PathReplay(batch=2, n_points=4096, k=16, first_scene=11)
    r2.f_dec = r.f_dec
    r2.F = r.F
    ref2 = r2.step().item()
    assert abs(loss2 - ref2) <= 1e-6 * abs(ref2)


def test_mm_replay_with_refinement_and_ignore_index():
    """BASELINE config 3 shape in small: 20 classes + ignored labels, DualMasks refinement before the loss."""
    from amcontrast3d_b200.replay import PathReplay
    r = PathReplay(batch=2, n_points=4096, k=12, num_classes=20, ignore_index=-100, refine=True, refine_k=8)
    loss = r.step()
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    for t in r.f_dec:
        assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0


def test_full_size_fps_chain_properties():
    """8 x 24000 -> 6000 -> 1500: one scene against the oracle bit for bit; all scenes: distinct indices, first
    pick 0, and — FPS of a cloud that is already in FPS order — picks in index order."""
    from amcontrast3d_b200.layers import furthest_point_sample
    xyz, _ = scenes.batch_of_scenes(8, 24000, "surface")
    p = torch.from_numpy(xyz).cuda()
    idx = furthest_point_sample(p, 6000)
    h = idx.cpu().numpy()
    assert (h[:, 0] == 0).all()
    for b in range(8):
        assert len(np.unique(h[b])) == 6000
    ridx, _ = oo.fps(xyz[3:4], 6000)
    assert np.array_equal(h[3:4], ridx)
    p1 = torch.gather(p, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    idx1 = furthest_point_sample(p1, 1500).cpu().numpy()
    assert np.array_equal(idx1, np.broadcast_to(np.arange(1500, dtype=np.int32), (8, 1500)))


def test_full_size_grouping_roundtrip_linearity():
    """config-2 sized grouping (8,128,6000)x(6000,32): gather is exact against torch, and the scatter-add is
    the adjoint of the gather: <G(f), g> == <f, G^T(g)> (checksum of checksums, size-independent)."""
    from amcontrast3d_b200.layers import ball_query, grouping_operation
    xyz, _ = scenes.batch_of_scenes(8, 6000, "surface", first_scene=20)
    p = torch.from_numpy(xyz).cuda()
    idx = ball_query(0.2, 32, p, p)
    f = torch.randn(8, 128, 6000, device="cuda", requires_grad=True)
    out = grouping_operation(f, idx)
    ref = torch.gather(f.detach(), 2, idx.reshape(8, 1, -1).expand(-1, 128, -1).long()).reshape(out.shape)
    assert torch.equal(out.detach(), ref)
    g = torch.randn_like(out)
    out.backward(g)
    terms = out.detach().double() * g.double()
    lhs = terms.sum()
    rhs = (f.detach().double() * f.grad.double()).sum()
    # the sum of ~2e8 zero-mean terms cancels to O(sqrt(N)): compare on that scale (FP32 scatter-add rounding)
    assert abs(lhs - rhs) <= 1e-5 * terms.pow(2).sum().sqrt()


def test_prefetch_pipeline_matches_unpipelined_batches():
    """The cross-step pipeline (FPS chain + first ball query of batch i+1 computed during step i) returns, one
    call late, exactly the losses of the unpipelined replay on the same batches — eager and as a CUDA graph."""
    from amcontrast3d_b200.replay import PathReplay
    batches = [scenes.batch_of_scenes(2, 4096, "surface", first_scene=s) for s in (0, 11, 23)]
    dev = [(torch.from_numpy(x).cuda(), torch.from_numpy(l).cuda()) for x, l in batches]
    ref = []
    for s in (0, 11, 23):
        r0 = PathReplay(batch=2, n_points=4096, k=16, first_scene=s, prefetch=False)
        ref.append(r0.step().item())
    r = PathReplay(batch=2, n_points=4096, k=16, first_scene=0, prefetch=True)
    got = [r.step(*dev[1]).item(), r.step(*dev[2]).item(), r.step(*dev[0]).item(), r.step(*dev[1]).item()]
    want = [ref[0], ref[1], ref[2], ref[0]]
    for g, w in zip(got, want):
        assert abs(g - w) <= 1e-6 * abs(w), (got, want)
    # the same through a captured graph.  Kernels do not run while a graph is being captured, so after
    # capture() the batch in flight is still the last one fed eagerly (batch 1)
    r.capture(warmup=1)
    pin = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(l).pin_memory()) for x, l in batches]
    g = [r.step_graph(*pin[2]).item(),         # -> loss of batch 1 (in flight), submits batch 2
         r.step_graph(*pin[0]).item(),         # -> loss of batch 2
         r.step_graph(*pin[1]).item(),         # -> loss of batch 0
         r.step_graph(*pin[1]).item()]         # -> loss of batch 1
    for got_i, want_i in zip(g, [ref[1], ref[2], ref[0], ref[1]]):
        assert abs(got_i - want_i) <= 1e-6 * abs(want_i), (g, ref)


def test_step_statistics_match_torch_counts():
    """the packed per-step statistics (selected points per stage, per-class tp / union / count of the stand-in
    prediction) written inside the step == the same quantities from torch ops (main_AA.py:461,496-507 reduce them)"""
    from amcontrast3d_b200.replay import PathReplay
    from amcontrast3d_b200.dist import PackedStats
    r = PathReplay(batch=2, n_points=4096, k=16, with_grouping=False)
    layout = PackedStats(13)
    r.stats_sink = torch.zeros(layout.size, device="cuda")
    loss = r.step()
    torch.cuda.synchronize()
    st = layout.unpack(r.stats_sink.double().cpu())
    geo = r._am_geometry
    assert st["n_selected"] == [int(g["stats"][0]) for g in geo]
    t = geo[0]["cls"].long()
    pred = t[geo[0]["knn_idx"][:, 1].long()]
    count = torch.bincount(t, minlength=13)
    tp = torch.bincount(t[pred == t], minlength=13)
    union = count + torch.bincount(pred, minlength=13) - tp
    assert st["count"].tolist() == count.cpu().tolist() and int(count.sum()) == 2 * 4096
    assert st["tp"].tolist() == tp.cpu().tolist()
    assert st["union"].tolist() == union.cpu().tolist()
    assert abs(st["loss_sum"] - loss.item()) <= 1e-6 * abs(loss.item())


def test_fused_conv_replay_runs_without_grouped_tensors():
    """PathReplay(fused_conv=True): every encoder grouping goes through the fused operator; the step runs eagerly and
    as a CUDA graph, gradients reach the stand-in features and the conv / BatchNorm parameters, and the largest
    tensor the step allocates is far smaller than a grouped tensor"""
    from amcontrast3d_b200.replay import PathReplay
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    r = PathReplay(batch=2, n_points=4096, k=16, fused_conv=True)
    loss = r.step()
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    for (l, i), (w, bn) in r.fparams.items():
        assert w.grad is not None and torch.isfinite(w.grad).all() and w.grad.abs().sum() > 0
        assert bn.weight.grad is not None and int(bn.num_batches_tracked) >= 1
    for t in r.F[:-1]:
        assert t.grad is not None and torch.isfinite(t.grad).all()
    # the plain replay's largest grouped tensor at this size is 2 x 128 x 1024 x 32 floats = 33.5 MB; the fused step
    # stays below the sum of the 19 grouped tensors by a wide margin
    peak = torch.cuda.max_memory_allocated() - base
    plain = PathReplay(batch=2, n_points=4096, k=16)
    torch.cuda.reset_peak_memory_stats()
    base2 = torch.cuda.memory_allocated()
    plain.step()
    torch.cuda.synchronize()
    assert peak < 0.6 * (torch.cuda.max_memory_allocated() - base2)
    # (bench.py replays this step as one CUDA graph in a fresh process: `fused_step` in its JSON line)
