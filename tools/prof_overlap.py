"""How much of the pipelined step is interference between its two halves?  Times, as CUDA graphs at config 2:
the whole pipelined step, the geometry half alone (FPS chain, ball queries, three_nn, loss geometry on the two
side streams) and the feature half alone (grouping / interpolation / loss forward + backward on static geometry)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from amcontrast3d_b200.replay import PathReplay


def timed(graph, reps=30):
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def capture(fn, warm=2):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


r = PathReplay(8, 24000, "cuda", 16, prefetch=True)
r.step()
r.step()
torch.cuda.synchronize()


def geometry_only(which=("enc", "aux")):
    main = torch.cuda.current_stream()
    r._geo.wait_stream(main)
    r._geo2.wait_stream(main)
    wl, wg = r.with_loss, r.with_grouping
    r.with_loss = wl and "aux" in which
    if "enc" not in which:
        r.with_grouping = False
    G = r._geometry(r.d_xyz_next, r.d_labels_next, r._geo, r._geo2)
    r.with_loss, r.with_grouping = wl, wg
    main.wait_stream(r._geo)
    main.wait_stream(r._geo2)
    return G


def feature_only():
    r.zero_grads()
    nlev = len(r.arch["blocks"])
    G = r._pf["cur"]
    p = [r.d_xyz] + G["p"]
    outs = []
    from amcontrast3d_b200.layers import three_interpolation
    for l in range(1, nlev):
        dp, fj = r.sa[l](p[l], p[l - 1], r.F[l - 1], idx=G["sa"][l])
        outs.append(fj)
        for i in range(r.arch["blocks"][l] - 1):
            dp, fj = r.la[l](p[l], p[l], r.F[l], idx=G["la"][l][i])
            outs.append(fj)
    for l in range(nlev - 1, 0, -1):
        outs.append(three_interpolation(p[l - 1], p[l], r.F[l], nn=G["nn3"][l]))
    r._am_geometry = G["am"]
    loss = r._loss_tail(p, r.d_labels)
    torch.autograd.backward(list(outs) + [loss], [r._grad_like(o) for o in outs] + [None])


g_feat = capture(feature_only)
t_feat = timed(g_feat)
g_geo = capture(geometry_only)
t_geo = timed(g_geo)
g_enc = capture(lambda: geometry_only(("enc",)))
t_enc = timed(g_enc)
g_aux = capture(lambda: geometry_only(("aux",)))
t_aux = timed(g_aux)
g_full = capture(r.step)
t_full = timed(g_full)
print(f"feature half alone      {t_feat:.3f} ms")
print(f"geometry half alone     {t_geo:.3f} ms   (encoder stream alone {t_enc:.3f}, FPS + loss stream alone {t_aux:.3f})")
print(f"whole pipelined step    {t_full:.3f} ms   -> interference {t_full - max(t_feat, t_geo):.3f} ms over the longer half")
