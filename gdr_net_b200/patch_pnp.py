"""Stand-alone Patch-PnP: `ConvPnPNet.forward(coor_feat, region, extents)` of the reference
(core/gdrn_modeling/models/conv_pnp_net.py:111-157) on the sm_100a kernels -- BASELINE.json configs[3]
("64x64x(3+64) correspondence maps -> R|t, batch 512").

    [B, 3|5, 64, 64] xyz (+2-D coords)  +  [B, 64, 64, 64] region attention  (+ extents [B, 3])
        -> gdrn_pnp_pack_input   NCHW fp32 -> NHWC 16-bit planes, 67 | 69 valid of 128 channels, xyz de-normalised
        -> 3 x (tcgen05 implicit-GEMM conv 3x3 s2 -> GroupNorm(32)+ReLU)        64 -> 32 -> 16 -> 8
        -> fc1 (8192 -> 1024, LeakyReLU 0.1) -> fc2 (1024 -> 256, LeakyReLU) -> fc_r | fc_t (one 9-column GEMM)
        -> rot [B, 6], t [B, 3]  (fp32)

Inference only (no autograd): inside `GDRN.forward` the same kernels run with saved activations and a hand-written
backward (engine.py).  The whole sequence is replayed as ONE CUDA graph per batch shape.  There is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import ops
from .capi import C
from .ops import PT, _stream


class PatchPnP:
    def __init__(self, pnp_net, precision: str = "fp32x3"):
        precision = {"mixed": "fp32x3", "fp16": "half", "bf16": "half"}.get(precision, precision)
        assert precision in ("half", "fp32x3"), precision
        C.load()
        self.net = pnp_net
        self.planes = 1 if precision == "half" else 2
        self.precision = precision
        self.wf = {}
        self._wkey = None
        self._graphs = {}
        self.use_cuda_graphs = True

    # ------------------------------------------------------------------ weights (re-packed only when a parameter changed)
    def _prepare_weights(self):
        n = self.net
        key = tuple((p.data_ptr(), p._version) for p in n.parameters())
        if key == self._wkey:
            return
        pl, wf = self.planes, self.wf
        dev = n.fc1.weight.device
        for ci in (0, 3, 6):
            wf[ci] = ops.pack_conv_fwd(n.features[ci].weight.detach(), pl, out=wf.get(ci), ipad=128)
        wf["fc1"] = ops.pack_linear(n.fc1.weight.detach(), pl, out=wf.get("fc1"), nhwc_from=(128, 8, 8))
        wf["fc2"] = ops.pack_linear(n.fc2.weight.detach(), pl, out=wf.get("fc2"))
        if "w_rt" not in wf:
            wf["w_rt"] = torch.empty(n.fc_r.out_features + 3, n.fc_r.in_features, device=dev)
            wf["b_rt"] = torch.empty(n.fc_r.out_features + 3, device=dev)
        with torch.no_grad():
            torch.cat([n.fc_r.weight, n.fc_t.weight], 0, out=wf["w_rt"])
            torch.cat([n.fc_r.bias, n.fc_t.bias], 0, out=wf["b_rt"])
        wf["fc_rt"] = ops.pack_linear(wf["w_rt"], pl, out=wf.get("fc_rt"))
        self._wkey = key
        self._graphs.clear()  # graphs hold the packed operands' addresses; buffers are reused, but stay safe

    # ------------------------------------------------------------------ the kernel sequence
    def _run(self, coor_feat, region, extents, pred):
        n, pl = self.net, self.planes
        B, c_feat = coor_feat.shape[0], coor_feat.shape[1]
        c_reg = 0 if region is None else region.shape[1]
        x = PT((B, 64, 64, 128), pl, device=coor_feat.device)
        C.gdrn_pnp_pack_input(coor_feat.data_ptr(), c_feat, ops.ptr(region), c_reg, ops.ptr(extents), x.hi_ptr, x.lo_ptr, B, 4096,
                              _stream())
        cur = x
        for ci, gi in ((0, 1), (3, 4), (6, 7)):
            u = ops.conv_fwd(cur, self.wf[ci], 128, 3, 3, 2, 1)
            gstats = torch.empty(B, 32, 2, device=coor_feat.device)
            cur = ops.gn_relu_fwd(u, n.features[gi].weight, n.features[gi].bias, gstats, G=n.features[gi].num_groups,
                                  eps=n.features[gi].eps)
        h1 = ops.gemm_fwd(cur.view(B, 8192), self.wf["fc1"], 1024, bias=n.fc1.bias, act=1)
        h2 = ops.gemm_fwd(h1, self.wf["fc2"], 256, bias=n.fc2.bias, act=1)
        ops.gemm_fwd(h2, self.wf["fc_rt"], n.fc_r.out_features + 3, out_f32=pred, bias=self.wf["b_rt"], ldc=16, want_planes=False)

    @torch.no_grad()
    def __call__(self, coor_feat: torch.Tensor, region: torch.Tensor | None = None, extents: torch.Tensor | None = None):
        if not coor_feat.is_cuda:
            raise RuntimeError("gdr_net_b200 Patch-PnP runs on CUDA (sm_100a) only; there is no CPU fallback")
        n = self.net
        coor_feat = coor_feat.float().contiguous()
        region = None if region is None else region.float().contiguous()
        extents = None if extents is None else extents.float().contiguous()
        B, c_feat, H, W = coor_feat.shape
        c_reg = 0 if region is None else region.shape[1]
        if (H, W) != (64, 64) or c_feat + c_reg != n.nIn:
            raise ValueError(f"Patch-PnP expects [B, {n.nIn} = coor + region channels, 64, 64]; got {c_feat} + {c_reg} at {H}x{W}")
        if c_feat in (3, 5) and extents is None:
            raise ValueError("extents are required to de-normalise the xyz channels (conv_pnp_net.py:120-122)")
        self._prepare_weights()
        nr = n.fc_r.out_features
        if not self.use_cuda_graphs:
            pred = torch.zeros(B, 16, device=coor_feat.device)
            self._run(coor_feat, region, extents, pred)
            return pred[:, :nr].clone(), pred[:, nr:nr + 3].clone()
        key = (B, c_feat, c_reg, extents is not None)
        g = self._graphs.get(key)
        if g is None:
            st = dict(coor=coor_feat.clone(), region=None if region is None else region.clone(),
                      ext=None if extents is None else extents.clone(), pred=torch.zeros(B, 16, device=coor_feat.device))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._run(st["coor"], st["region"], st["ext"], st["pred"])
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            st["graph"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st["graph"]):
                self._run(st["coor"], st["region"], st["ext"], st["pred"])
            g = self._graphs[key] = st
        g["coor"].copy_(coor_feat, non_blocking=True)
        if region is not None:
            g["region"].copy_(region, non_blocking=True)
        if extents is not None:
            g["ext"].copy_(extents, non_blocking=True)
        g["graph"].replay()
        return g["pred"][:, :nr].clone(), g["pred"][:, nr:nr + 3].clone()
