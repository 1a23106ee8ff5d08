"""Path replay: one training iteration's worth of hot-path calls, in the reference's call order.

What a PointNeXt-XL / AMContrast3D step executes on this path,
with random tensors standing in for the MLP outputs so that no cuDNN/cuBLAS work is timed:

  encoder, level l = 1..4   furthest_point_sample(p_{l-1}, n_l) ; gather p_l
                            QueryAndGroup(r_l, 32)(p_l, p_{l-1}, F_{l-1})          SetAbstraction
                            (blocks_l - 1) x QueryAndGroup(2 r_l, 32)(p_l, p_l, F_l)   InvResMLP
  decoder, level l = 4..1   three_interpolation(p_{l-1}, p_l, F_l)                 FeaturePropogation
  criterion                 ContrastHead over the 4 stages {p_s (M_s,3), f_s (M_s,D_s), offset=[M_s]}
  backward                  loss -> grad f_s ; grouping_operation scatter-add for the 19 feature
                            groupings ; three_interpolate gradients
  (AMContrast3D++)          RefinementMethod.DualMasks per decoder stage, forward + backward

Every call goes through the reference-facing Python operators (amcontrast3d_b200.layers,
.AMContrast3D), i.e. through the public API a user of the reference would call.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import scenes
from .AMContrast3D import ContrastHead, RefinementMethod
from .layers import QueryAndGroup, furthest_point_sample, three_interpolation

# PointNeXt-XL as configured by cfgs/s3dis/AMContrast3D-AA.yaml:34-44
XL = dict(blocks=(1, 4, 7, 4, 4), strides=(1, 4, 4, 4, 4), width=64, radius=0.1, nsample=32)


def aa_args(nsample=16, **kw):
    """ambiguity_args of cfgs/s3dis/AMContrast3D-AA.yaml:6-30 (nsample per BASELINE config)."""
    d = dict(nsample=nsample, ccbeta=0.04, cctype="Method2", temperature=0.3, supervisedCL="Method1", db="-m",
             margin="adaptive", mu=-1, nu=0.5, stages="up", stages_num=4, vis=False, w1=0.1, w2=0.9, w3=0.01)
    d.update(kw)
    return SimpleNamespace(**d)


class PathReplay:
    """Holds the synthetic inputs of one data-parallel unit (B scenes flattened into one segment
    by the encoder, pointnext_AA.py:458-462) and replays the hot path over them."""

    def __init__(self, batch=8, n_points=24000, device="cuda", k=16, num_classes=13, ignore_index=None,
                 kind="surface", rank=0, first_scene=0, arch=XL, refine=False, refine_k=12, seed=0,
                 with_grouping=True, with_loss=True, geometry_stream=True, prefetch=False, loss_args=None,
                 fused_conv=False):
        self.B, self.N, self.device = batch, n_points, torch.device(device)
        self.num_classes, self.ignore_index = num_classes, ignore_index
        self.args = aa_args(k, **(loss_args or {}))      # e.g. ScanNet / MM: temperature=0.5, nu=0.6
        self.arch = arch
        self.refine, self.refine_k = refine, refine_k
        self.with_grouping, self.with_loss = with_grouping, with_loss
        # FPS, ball_query and three_nn depend on coordinates only.  With geometry_stream they run on a side
        # stream that works ahead of the feature path (in a real step: ahead of the MLPs as well): the
        # latency-bound FPS clusters and the search kernels overlap the HBM-bound grouping kernels.  Same
        # operator calls, same results; only the issue order and the stream differ.
        self.geometry_stream = geometry_stream and torch.cuda.is_available()
        self._geo = None
        self._geo2 = None
        self._am_geometry = None
        # prefetch: a software pipeline across steps.  The geometry of a batch (FPS chain, ball queries,
        # three_nn, the loss's labels / kNN / ambiguity) depends on its coordinates and labels only, so it is
        # computed one step EARLY, on side streams, while the previous batch's feature path and backward keep
        # HBM busy — the way a data loader prefetches.  Every step still runs one full geometry pass and one
        # full feature pass; the feature path no longer waits for any search.
        self.prefetch = prefetch and self.geometry_stream
        self._pf = None           # {'cur': static geometry of the batch in flight, 'next': this step's}
        xyz, labels = scenes.batch_of_scenes(batch, n_points, kind, rank=rank, first_scene=first_scene,
                                             num_classes=num_classes,
                                             ignore_fraction=0.05 if ignore_index is not None else 0.0)
        # host side of the step: pinned xyz + labels (what the trainer's H2D copy moves, main_AA.py:377-379)
        self.h_xyz = torch.from_numpy(xyz).pin_memory() if torch.cuda.is_available() else torch.from_numpy(xyz)
        self.h_labels = torch.from_numpy(labels).pin_memory() if torch.cuda.is_available() else torch.from_numpy(labels)
        self.h2d_bytes = self.h_xyz.numel() * 4 + self.h_labels.numel() * 8
        self.d_xyz = self.h_xyz.to(self.device)
        self.d_labels = self.h_labels.to(self.device)

        nlev = len(arch["blocks"])
        self.n = [n_points]
        self.C = [arch["width"]]
        for l in range(1, nlev):
            self.n.append(self.n[-1] // arch["strides"][l])
            self.C.append(self.C[-1] * 2)
        g = torch.Generator(device=self.device)
        g.manual_seed(1234 + seed + 7919 * rank)
        # F_l: stand-ins for the MLP outputs at each level, (B, C_l, n_l), require grad
        self.F = [torch.randn((batch, self.C[l], self.n[l]), device=self.device, generator=g).requires_grad_(True)
                  for l in range(nlev)]
        # decoder features feeding the loss, (M_s, D_s) row-major, require grad
        self.f_dec = [torch.randn((batch * self.n[s], self.C[s]), device=self.device, generator=g).requires_grad_(True)
                      for s in range(4)]
        r = arch["radius"]
        self.sa = [None] + [QueryAndGroup(r * 2 ** (l - 1), arch["nsample"], normalize_dp=True) for l in range(1, nlev)]
        self.la = [None] + [QueryAndGroup(r * 2 ** l, arch["nsample"], normalize_dp=True) for l in range(1, nlev)]
        self.head = ContrastHead()
        # fused_conv: every grouping of the encoder is replaced by the fused grouping -> 1x1 conv -> BatchNorm -> ReLU ->
        # max operator (layers/fused.py) with that layer's conv / BatchNorm parameters — what a PointNeXt step executes
        # when compat tier 4 is installed.  The (B, 3+C, M, 32) grouped tensors then never exist; the step now contains
        # the 19 convolutions (865 GFLOP forward at config 2), which the plain replay leaves to the model.
        self.fused_conv = fused_conv
        if fused_conv:
            import torch.nn as nn
            self.fparams = {}
            for l in range(1, nlev):
                for i in range(arch["blocks"][l]):                      # i = 0: SetAbstraction, i >= 1: LocalAggregation
                    cin = self.C[l - 1] if i == 0 else self.C[l]
                    w = (torch.randn((self.C[l], cin + 3), device=self.device, generator=g) / (cin + 3) ** 0.5).requires_grad_(True)
                    self.fparams[(l, i)] = (w, nn.BatchNorm2d(self.C[l]).to(self.device))
        # synthetic upstream gradients: one buffer, viewed per output
        self._gbuf = None
        self.points_per_step = batch * n_points

    # ------------------------------------------------------------------------------------
    def _grad_like(self, t):
        n = t.numel()
        if self._gbuf is None or self._gbuf.numel() < n:
            self._gbuf = torch.randn(n, device=self.device)
        return self._gbuf[:n].view(t.shape)

    def zero_grads(self):
        for t in self.F + self.f_dec:
            t.grad = None
        if self.fused_conv:
            for w, bn in self.fparams.values():
                w.grad = None
                bn.weight.grad = None
                bn.bias.grad = None

    def _aggregate(self, l, i, grouper, q, s, feats, idx=None):
        """One neighbourhood aggregation of the encoder: the grouped features (what the reference's modules feed their
        convolution), or — fused_conv — the output of the fused grouping + conv + BatchNorm + ReLU + max operator."""
        if not self.fused_conv:
            return grouper(q, s, feats, idx=idx)[1]
        from .layers import ball_query
        from .layers.fused import fused_group_conv_bn_relu_max
        if idx is None:
            idx = ball_query(grouper.radius, grouper.nsample, s, q)
        w, bn = self.fparams[(l, i)]
        return fused_group_conv_bn_relu_max(q, s, feats, idx, w, bn, grouper.radius, True, "tf32")

    def forward(self, xyz=None, labels=None):
        """-> (loss, outputs needing a synthetic upstream gradient)"""
        if self.prefetch:
            return self._forward_prefetch(xyz, labels)
        p0 = self.d_xyz if xyz is None else xyz
        labels = self.d_labels if labels is None else labels
        arch = self.arch
        p = [p0]
        outs = []
        nlev = len(arch["blocks"])
        if self.geometry_stream:
            from .layers import ball_query, three_nn
            main = torch.cuda.current_stream(self.device)
            if self._geo is None:
                self._geo = torch.cuda.Stream(device=self.device)
            geo = self._geo
            geo.wait_stream(main)                                   # fork (inputs are ready on `main`)
            ev_lvl, ev_up, bq_sa, bq_la, nn3 = [None], [None] * nlev, [None], [None], [None] * nlev
            ev_pts = [None] * nlev
            with torch.cuda.stream(geo):
                for l in range(1, nlev):
                    idx = furthest_point_sample(p[l - 1], self.n[l]).long()
                    p.append(torch.gather(p[l - 1], 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous())
                    ev_pts[l] = torch.cuda.Event()
                    ev_pts[l].record(geo)
                    sa_idx, la_idx = None, []
                    if self.with_grouping:
                        sa_idx = ball_query(self.sa[l].radius, self.sa[l].nsample, p[l - 1], p[l])
                        la_idx = [ball_query(self.la[l].radius, self.la[l].nsample, p[l], p[l])
                                  for _ in range(arch["blocks"][l] - 1)]
                    bq_sa.append(sa_idx)
                    bq_la.append(la_idx)
                    ev = torch.cuda.Event()
                    ev.record(geo)
                    ev_lvl.append(ev)
                for l in range(nlev - 1, 0, -1):
                    nn3[l] = three_nn(p[l - 1], p[l])
                    ev_up[l] = torch.cuda.Event()
                    ev_up[l].record(geo)
            # the loss's own geometry (stage labels, kNN, posmask, ambiguity: coordinates + labels only) on a
            # second side stream: stage 0 needs nothing but the input cloud, so its 192 000-point kNN runs
            # while the first FPS — a latency-bound chain that leaves most of the GPU idle — is in flight
            if self.with_loss:
                if self._geo2 is None:
                    self._geo2 = torch.cuda.Stream(device=self.device)
                geo2 = self._geo2
                geo2.wait_stream(main)
                with torch.cuda.stream(geo2):
                    pts, self._am_geometry = [], []
                    for s in range(4):
                        if s > 0:
                            geo2.wait_event(ev_pts[s])
                        pts.append({"p_out": p[s].reshape(-1, 3), "offset": self._offsets[s]})
                        sl = {"down": pts, "up": pts}
                        self._am_geometry.append(self.head.precompute_geometry(
                            s, sl, labels.reshape(-1), self.num_classes, self.ignore_index, self.args))
            for l in range(1, nlev):
                main.wait_event(ev_lvl[l])
                if self.with_grouping:
                    outs.append(self._aggregate(l, 0, self.sa[l], p[l], p[l - 1], self.F[l - 1], bq_sa[l]))
                    for i in range(arch["blocks"][l] - 1):
                        outs.append(self._aggregate(l, i + 1, self.la[l], p[l], p[l], self.F[l], bq_la[l][i]))
            for l in range(nlev - 1, 0, -1):
                main.wait_event(ev_up[l])
                outs.append(three_interpolation(p[l - 1], p[l], self.F[l], nn=nn3[l]))
            main.wait_stream(geo)                                   # join
            if self.with_loss:
                main.wait_stream(self._geo2)
        else:
            self._am_geometry = None
            for l in range(1, nlev):
                idx = furthest_point_sample(p[l - 1], self.n[l]).long()
                p.append(torch.gather(p[l - 1], 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous())
                if self.with_grouping:
                    outs.append(self._aggregate(l, 0, self.sa[l], p[l], p[l - 1], self.F[l - 1]))
                    for i in range(arch["blocks"][l] - 1):
                        outs.append(self._aggregate(l, i + 1, self.la[l], p[l], p[l], self.F[l]))
            for l in range(nlev - 1, 0, -1):
                outs.append(three_interpolation(p[l - 1], p[l], self.F[l]))
        return self._loss_tail(p, labels), outs

    def _loss_tail(self, p, labels):
        loss = None
        if self.with_loss:
            feats = self.f_dec
            if self.refine:
                feats = []
                for s in range(4):
                    B, n_s, D = self.B, self.n[s], self.C[s]
                    f_bdn = self.f_dec[s].view(B, n_s, D).transpose(1, 2).contiguous()
                    a = self._apm[s]
                    f_ref, _ = RefinementMethod({}, p[s], f_bdn, a, s, B, self.refine_k, "MIN", 1.0, 0.9, 0.4).DualMasks()
                    feats.append(f_ref.transpose(1, 2).reshape(B * n_s, D))
            down = [{"p_out": p[s].reshape(-1, 3), "f_out": feats[s],
                     "offset": self._offsets[s]} for s in range(4)]
            stage_list = {"inputs": None, "down": down, "up": down}
            if self._am_geometry is not None:
                stage_list["am_geometry"] = self._am_geometry
            loss, a_cat, _ = self.head(None, labels.reshape(-1), stage_list, self.num_classes, self.ignore_index,
                                       self.args)
        return loss

    # ------------------------------------------------------------------------------------
    def _fps_chain(self, p0):
        q = [p0]
        for l in range(1, len(self.arch["blocks"])):
            idx = furthest_point_sample(q[l - 1], self.n[l]).long()
            q.append(torch.gather(q[l - 1], 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous())
        return q

    def _geometry(self, xyz, labels, enc, aux):
        """Everything of one batch that depends on coordinates and labels only — FPS chain, all ball queries,
        the three_nn searches and the loss's label / kNN / ambiguity geometry — issued on the streams `enc`
        (encoder geometry) and `aux` (loss geometry, which only waits for the points of its stage)."""
        from .layers import ball_query, three_nn
        arch, nlev = self.arch, len(self.arch["blocks"])
        G = {"p": [], "sa": [None], "la": [None], "nn3": [None] * nlev, "am": None}
        ev_pts = [None] * nlev
        with torch.cuda.stream(enc):
            q = [xyz]
            for l in range(1, nlev):
                idx = furthest_point_sample(q[l - 1], self.n[l]).long()
                q.append(torch.gather(q[l - 1], 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous())
                ev_pts[l] = torch.cuda.Event()
                ev_pts[l].record(enc)
            G["p"] = q[1:]
            for l in range(1, nlev):
                sa_idx, la_idx = None, []
                if self.with_grouping:
                    sa_idx = ball_query(self.sa[l].radius, self.sa[l].nsample, q[l - 1], q[l])
                    la_idx = [ball_query(self.la[l].radius, self.la[l].nsample, q[l], q[l])
                              for _ in range(arch["blocks"][l] - 1)]
                G["sa"].append(sa_idx)
                G["la"].append(la_idx)
            for l in range(nlev - 1, 0, -1):
                G["nn3"][l] = three_nn(q[l - 1], q[l])
        if self.with_loss:
            with torch.cuda.stream(aux):
                pts, am = [], []
                sl = {"down": pts, "up": pts}
                for s in range(4):
                    if s > 0:
                        aux.wait_event(ev_pts[s])
                    pts.append({"p_out": q[s].reshape(-1, 3), "offset": self._offsets[s]})
                    am.append(self.head.precompute_geometry(s, sl, labels.reshape(-1), self.num_classes,
                                                            self.ignore_index, self.args))
                G["am"] = am
        return G

    _AM_KEYS = ("knn_idx", "posbits", "cnt", "a", "stats", "cls", "order")

    def _copy_geometry(self, dst, src):
        """dst <- src for every tensor a later step reads (the static copy the captured graph is bound to):
        ~60 small tensors, moved with one multi-tensor copy per dtype instead of 60 launches."""
        pairs = list(zip(dst["p"], src["p"]))
        for l in range(1, len(dst["sa"])):
            if dst["sa"][l] is not None:
                pairs.append((dst["sa"][l], src["sa"][l]))
            pairs += list(zip(dst["la"][l], src["la"][l]))
        for d, s_ in zip(dst["nn3"], src["nn3"]):
            if d is not None:
                pairs += [(d[0], s_[0]), (d[1], s_[1])]
        if dst["am"] is not None:
            for d, s_ in zip(dst["am"], src["am"]):
                pairs += [(d[k], s_[k]) for k in self._AM_KEYS if d.get(k) is not None]
        pairs += [(self.d_xyz, self.d_xyz_next), (self.d_labels, self.d_labels_next)]
        by_dtype = {}
        for d, s_ in pairs:
            by_dtype.setdefault(d.dtype, ([], []))
            by_dtype[d.dtype][0].append(d)
            by_dtype[d.dtype][1].append(s_)
        for dl, sl in by_dtype.values():
            torch._foreach_copy_(dl, sl)

    def _forward_prefetch(self, xyz=None, labels=None):
        """Pipelined schedule: `xyz` / `labels` (if given) are the NEXT batch, whose whole geometry is computed
        on side streams during this call; the batch whose geometry was computed during the previous call goes
        through the feature path now, without waiting for any search."""
        arch, nlev = self.arch, len(self.arch["blocks"])
        main = torch.cuda.current_stream(self.device)
        if self._pf is None:
            # prologue (eager, once): streams, buffers for the next batch, the current batch's geometry
            for name in ("_geo", "_geo2"):
                setattr(self, name, torch.cuda.Stream(device=self.device))
            self.d_xyz_next, self.d_labels_next = self.d_xyz.clone(), self.d_labels.clone()
            for st in (self._geo, self._geo2):
                st.wait_stream(main)
            self._pf = {"cur": self._geometry(self.d_xyz, self.d_labels, self._geo, self._geo2), "next": None}
            main.wait_stream(self._geo)
            main.wait_stream(self._geo2)
        if xyz is not None:
            self.d_xyz_next.copy_(xyz, non_blocking=True)
        if labels is not None:
            self.d_labels_next.copy_(labels, non_blocking=True)
        geo, geo2 = self._geo, self._geo2
        geo.wait_stream(main)                                       # fork
        geo2.wait_stream(main)
        self._pf["next"] = self._geometry(self.d_xyz_next, self.d_labels_next, geo, geo2)
        # the current batch: every index it needs exists already
        G = self._pf["cur"]
        p = [self.d_xyz] + G["p"]
        outs = []
        for l in range(1, nlev):
            if self.with_grouping:
                outs.append(self._aggregate(l, 0, self.sa[l], p[l], p[l - 1], self.F[l - 1], G["sa"][l]))
                for i in range(arch["blocks"][l] - 1):
                    outs.append(self._aggregate(l, i + 1, self.la[l], p[l], p[l], self.F[l], G["la"][l][i]))
        for l in range(nlev - 1, 0, -1):
            outs.append(three_interpolation(p[l - 1], p[l], self.F[l], nn=G["nn3"][l]))
        self._am_geometry = G["am"]
        return self._loss_tail(p, self.d_labels), outs

    def _rotate_prefetch(self):
        """End of a pipelined step: the geometry computed for the next batch becomes the current batch's."""
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self._geo)                                 # join
        main.wait_stream(self._geo2)
        self._copy_geometry(self._pf["cur"], self._pf["next"])      # includes xyz / labels of the next batch
        self._pf["next"] = None

    @property
    def _offsets(self):
        if not hasattr(self, "_off"):
            self._off = [torch.tensor([self.B * self.n[s]], dtype=torch.int32, device=self.device) for s in range(4)]
        return self._off

    @property
    def _apm(self):
        if not hasattr(self, "_apm_a"):
            g = torch.Generator(device=self.device)
            g.manual_seed(99)
            self._apm_a = [torch.rand((self.B, 1, self.n[s]), device=self.device, generator=g) for s in range(4)]
        return self._apm_a

    def step(self, xyz=None, labels=None):
        """forward + backward of the path; returns the loss tensor (device scalar)"""
        self.zero_grads()
        loss, outs = self.forward(xyz, labels)
        tensors = list(outs)
        grads = [self._grad_like(o) for o in outs]
        if loss is not None:
            tensors.append(loss)
            grads.append(None)
        torch.autograd.backward(tensors, grads)
        if self.stats_sink is not None and loss is not None:
            self._write_stats(loss)
        if self.prefetch:
            self._rotate_prefetch()
        return loss

    # the step's statistics, as the trainer reduces them across ranks (main_AA.py:461,496-507): packed into ONE
    # float32 vector [loss_sum, CE, AM, n_selected[4], tp[ncls], union[ncls], count[ncls]] that the caller owns
    # (bench.py: the tail of the last gradient bucket, so that no separate small all-reduce is exposed after the
    # step).  Static shapes, no host read: part of the captured graph.
    stats_sink = None

    def _write_stats(self, loss):
        ncls = self.num_classes + (1 if self.ignore_index is not None else 0)
        geo = self._am_geometry
        n_sel = torch.stack([g["stats"][0] for g in geo]).float()                 # points the loss selected, per stage
        g0 = geo[0]
        cls, knn = g0["cls"], g0["knn_idx"]                     # stage-0 labels (ignored -> extra class); kNN rows
        counts = torch.empty(3 * ncls, dtype=torch.float32, device=cls.device)
        from . import _capi
        with _capi.guard(cls):                                  # prediction = the nearest neighbour's label (column 1)
            _capi.call("amc3d_class_counts", int(cls.shape[0]), ncls, _capi.ptr(cls), 0, knn.data_ptr() + 4,
                       int(knn.shape[1]), _capi.ptr(counts), _capi.stream(cls))
        tp, npred, count = counts[:ncls], counts[ncls:2 * ncls], counts[2 * ncls:]
        union = count + npred - tp
        lf = loss.detach().float().reshape(1)
        vec = torch.cat([lf, torch.zeros_like(lf), lf, n_sel, tp[:self.num_classes], union[:self.num_classes],
                         count[:self.num_classes]])
        self.stats_sink.copy_(vec)

    def step_e2e(self):
        """The same step driven from HOST buffers: pinned xyz/labels -> device, step, loss -> host."""
        xyz = self.h_xyz.to(self.device, non_blocking=True)
        labels = self.h_labels.to(self.device, non_blocking=True)
        loss = self.step(xyz, labels)
        return float(loss.item()) if loss is not None else 0.0

    # ------------------------------------------------------------------------------------
    # CUDA-graph replay.  Every shape on the path is static (B, N and the level sizes are fixed by
    # the config) and nothing on it reads a device value on the host, so one whole step — the ~170
    # C-ABI launches plus the torch glue of the operator wrappers and autograd — records into a
    # single graph.  Replaying it removes the per-call Python / launch cost (about 20 ms per step
    # at config 2, more than the kernels themselves) while executing exactly the same kernels.
    # ------------------------------------------------------------------------------------
    def capture(self, warmup: int = 3):
        """Record step() into a CUDA graph over static input buffers (self.d_xyz / self.d_labels).
        Returns self; afterwards step_graph() replays it."""
        from . import _capi
        assert _capi.PROFILE is None, "do not capture while per-call profiling is on"
        # the leaves' AccumulateGrad nodes were created on the default stream by earlier eager steps; the
        # capture stream differs by construction, which is harmless here (everything is ordered by the
        # capture) — silence torch's per-step warning about it
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):            # warm-up off the default stream, as torch requires
            for _ in range(max(1, warmup)):
                self.step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.zero_grads()
        l0 = _capi.LAUNCHES
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._graph_loss = self.step()
        self.graph_launches = _capi.LAUNCHES - l0
        return self

    def step_graph(self, h_xyz=None, h_labels=None):
        """Replay the captured step.  With host tensors given, they are first copied (async, from
        pinned memory) into the static device buffers the graph reads.  Returns the loss tensor."""
        if h_xyz is not None:
            (self.d_xyz_next if self.prefetch else self.d_xyz).copy_(h_xyz, non_blocking=True)
        if h_labels is not None:
            (self.d_labels_next if self.prefetch else self.d_labels).copy_(h_labels, non_blocking=True)
        self._graph.replay()
        return self._graph_loss
