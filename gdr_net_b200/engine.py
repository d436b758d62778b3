"""Forward / backward orchestration of the GDR-Net hot path over the C ABI (libgdrn_b200.so).

Mirrors reference `GDRN.forward` (core/gdrn_modeling/models/GDRN.py:83-306), `gdrn_loss` (:308-521) and the
autograd backward of both, but every tensor op is one of our sm_100a kernels (gdr_net_b200/csrc).  PyTorch is
used for device memory, streams and the autograd/DDP boundary only.

precision = "half"   : activations / gradients are 16-bit NHWC tensors (fp16 by default, bf16 if the library is built with
                       GDRN_STORE_F16=0), one tcgen05 pass per k-step.  "fp16" / "bf16" are accepted as aliases.
precision = "fp32x3" : (hi, lo) 16-bit planes (22-bit operands with fp16), three tcgen05 passes in forward AND backward
precision = "mixed"  : the fp32x3 forward (every output / loss inside the 1e-3 parity bound, identical ReLU / max-pool / L1
                       decisions) with the single-plane backward of "half": the backward is linear given the forward's
                       masks, so 11-bit operands there cost ~1e-3 on the gradients (what cuDNN's TF32 backward gives the
                       reference) instead of flipping decisions.  Saved activations are read through their hi plane.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .capi import C
from .ops import PT, _stream

import os

_DGRAD_ZERO_INSERT = os.environ.get("GDRN_DGRAD_ZERO_INSERT") == "1"  # A/B switch: the first (zero-insert + s1 conv) s2 dgrad

# A/B switch: recompute the ReLU mask of plain conv-BN-ReLU layers from u instead of reading y back in BN backward.
# MEASURED (same box, graphed step): 11.39-11.47 ms with the recompute vs 11.11-11.33 ms reading y -- 2 B/element less HBM
# traffic does not pay for the extra per-element FMA + compare + shared-memory constants, so it is off.
_BN_MASK_FROM_U = os.environ.get("GDRN_BN_MASK_FROM_U") == "1"
# A/B switch: ConvTranspose2d forward / wgrad by output-parity phases over the un-dilated input (default) instead of a
# stride-1 conv over the zero-inserted input (4x the MACs + a zero_insert pass)
_DECONV_PHASES = os.environ.get("GDRN_DECONV_PHASES", "1") == "1"
# A/B switch: BatchNorm forward emits the ReLU mask as a bitmap that backward reads instead of the activation (default on)
_BN_BITMASK = os.environ.get("GDRN_BN_BITMASK", "1") == "1"

LOSS_NAMES = ["loss_coor_x", "loss_coor_y", "loss_coor_z", "loss_mask", "loss_region", "loss_PM_R", "loss_centroid", "loss_z"]
HEAD_CONVS = [(3, 4, False), (6, 7, False), (10, 11, True), (13, 14, False), (17, 18, True), (20, 21, False)]


class _BN:
    """Per-BatchNorm scratch: scale/shift (forward), mean/invstd (backward), statistic slices."""

    def __init__(self, C_, dev):
        self.scale = torch.empty(C_, device=dev)
        self.shift = torch.empty(C_, device=dev)
        self.mean = torch.empty(C_, device=dev)
        self.invstd = torch.empty(C_, device=dev)
        self.sums = torch.empty(2, C_, device=dev)
        self.stats = None  # view into Engine.stats_all


class SymTable:
    """Device-resident table of symmetry rotations (reference `sym_infos`: per-object [K,3,3] model-to-model transforms,
    core/utils/pose_utils.py:457-482, built per dataset by misc.get_symmetry_transformations).  Each distinct object is
    uploaded ONCE; a step only sends B (first row, count) pairs, so the symmetric PM loss needs no per-step packing of
    the matrices, no host loop over candidates and is CUDA-graph capturable (fixed table address, fixed capacity)."""

    CAPACITY = 16384  # rows (YCB-V: 21 objects x <= 628 discretised transforms fits several times over)

    def __init__(self, dev):
        self.dev = dev
        self.table = torch.zeros(self.CAPACITY, 3, 3, device=dev)
        self.rows = 0
        self._by_id = {}     # id(obj) -> (obj, first, count): fast path for the per-dataset arrays reused every step
        self._by_bytes = {}  # content -> (first, count)

    def _register(self, s):
        import numpy as np

        arr = s.detach().cpu().numpy() if isinstance(s, torch.Tensor) else np.asarray(s)
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1, 3, 3)
        key = arr.tobytes()
        ent = self._by_bytes.get(key)
        if ent is None:
            k = arr.shape[0]
            if self.rows + k > self.CAPACITY:
                raise RuntimeError(f"symmetry table full ({self.rows} + {k} > {self.CAPACITY} rows)")
            self.table[self.rows:self.rows + k].copy_(torch.from_numpy(arr))
            ent = (self.rows, k)
            self.rows += k
            self._by_bytes[key] = ent
        return ent

    def indices(self, sym_infos) -> torch.Tensor:
        """list[B] of None | [K,3,3] -> int32 [B,2] (first row, count) on the device (one small async H2D copy)."""
        out = torch.zeros(len(sym_infos), 2, dtype=torch.int32).pin_memory()
        for b, s in enumerate(sym_infos):
            if s is None:
                continue
            hit = self._by_id.get(id(s))
            if hit is None or hit[0] is not s:
                first, count = self._register(s)
                if len(self._by_id) > 4096:
                    self._by_id.clear()
                self._by_id[id(s)] = (s, first, count)
            else:
                first, count = hit[1], hit[2]
            out[b, 0], out[b, 1] = first, count
        return out.to(self.dev, non_blocking=True)


class Engine:
    def __init__(self, model, precision: str = "half"):
        precision = {"bf16": "half", "fp16": "half"}.get(precision, precision)
        assert precision in ("half", "fp32x3", "mixed"), precision
        C.load()  # fail loudly if the CUDA library is missing
        self.model = model
        self.precision = precision
        self.planes = 1 if precision == "half" else 2      # forward activations / forward weight operands
        self.bplanes = 2 if precision == "fp32x3" else 1   # activation gradients / dgrad weight operands
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("gdr_net_b200 engine needs CUDA parameters (model.to('cuda')); no CPU fallback exists")
        self.ws = ops.Workspace(self.dev)
        self.sym_table = SymTable(self.dev)
        # deterministic reductions (bit-identical run to run): BatchNorm batch statistics and backward sums in two ordered
        # stages instead of fp32 atomics.  Costs one extra read of every pre-BN tensor in forward; a debugging / testing mode.
        self.wgrad_side_stream = os.environ.get("GDRN_WGRAD_STREAM", "1") == "1"
        self._side_stream, self._side_keep = None, []
        self.deterministic = os.environ.get("GDRN_DETERMINISTIC") == "1"
        self._det_ws = None
        self.fold_eval = os.environ.get("GDRN_NO_FOLD_EVAL") != "1"  # A/B switch: eval forward through the unfused train-path kernels
        self.with_2d = int(model.pnp_net.features[0].in_channels == 69)  # 69 = xyz + coord2d + regions, 67 without coords
        assert model.pnp_net.features[0].in_channels in (67, 69), model.pnp_net.features[0].in_channels
        self.named_params = list(model.named_parameters())
        self.bn: Dict[str, _BN] = {}
        total = 0
        self._bn_mods = {}
        for name, m in model.named_modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                self.bn[name] = _BN(m.num_features, self.dev)
                self._bn_mods[name] = m
                total += 2 * m.num_features
        self.stats_all = torch.zeros(total, device=self.dev)
        self.bwd_sums_all = torch.zeros(total, device=self.dev)  # BN-backward accumulators: one memset per backward
        off = 0
        for name, m in self._bn_mods.items():
            n = 2 * m.num_features
            self.bn[name].stats = self.stats_all[off:off + n]
            self.bn[name].sums = self.bwd_sums_all[off:off + n].view(2, m.num_features)
            off += n
        # flat gradient buffer (one NCCL-friendly allocation); per-parameter views
        self.flat_grad = torch.zeros(sum(p.numel() for _, p in self.named_params), device=self.dev)
        self.grads: Dict[str, torch.Tensor] = {}
        off = 0
        for name, p in self.named_params:
            self.grads[name] = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.saved = None
        self.wf: Dict[str, PT] = {}
        self.wd: Dict[str, PT] = {}
        self.grad_hook = None  # optional callable(engine) invoked at bucket boundaries during backward (DDP overlap)
        self.use_cuda_graphs = False  # module API: replay forward / backward CUDA graphs (fixed shapes, no symmetric PM)
        self._graphed = None
        # fp16 planes: activation gradients are carried with a static power-of-two loss scale (exactly removed from the
        # fp32 parameter gradients at the end of backward), like the GradScaler of the reference's AMP configs
        self.storage_name = ops.storage_format()[2]
        self.grad_scale = 1024.0 if self.storage_name == "fp16" else 1.0
        from .dist import SEGMENTS, segment_bounds

        self._seg = segment_bounds(self.named_params, SEGMENTS)

    # ------------------------------------------------------------------------------------------ weights
    def _conv_modules(self):
        m = self.model
        out = [("backbone.conv1", m.backbone.conv1)]
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(m.backbone, f"layer{li}")):
                p = f"backbone.layer{li}.{bi}"
                out.append((p + ".conv1", blk.conv1))
                out.append((p + ".conv2", blk.conv2))
                if blk.downsample is not None:
                    out.append((p + ".downsample.0", blk.downsample[0]))
        for ci, _bi, _u in HEAD_CONVS:
            out.append((f"rot_head_net.features.{ci}", m.rot_head_net.features[ci]))
        out.append(("rot_head_net.features.23", m.rot_head_net.features[23]))
        for ci in (0, 3, 6):
            out.append((f"pnp_net.features.{ci}", m.pnp_net.features[ci]))
        return out

    def prepare_weights(self, need_dgrad: bool):
        """fp32 parameters -> K-major bf16 (hi, lo) GEMM operands, re-done every step (weights change) by ONE batched
        launch over a device-resident job table (rebuilt only if a parameter / buffer pointer moved)."""
        import numpy as np

        key = (need_dgrad, tuple(p.data_ptr() for _, p in self.named_params))
        tables = self.__dict__.setdefault("_pack_tables", {})
        tab = tables.get(key)
        if tab is None:
            # One persistent job table per (need_dgrad, parameter addresses).  Tables are NEVER freed or overwritten: a
            # captured CUDA graph holds the raw device pointer of the table it was captured with, and an eager call with
            # the other need_dgrad value (train -> eval -> train) must not invalidate it.
            ops._pack_recorder = []
            try:
                self._prepare_weights_calls(need_dgrad)
                self._pack_stem()
                jobs = ops._pack_recorder
            finally:
                ops._pack_recorder = None
            dt = np.dtype([("src", "<u8"), ("hi", "<u8"), ("lo", "<u8"), ("so", "<i8"), ("si", "<i8"), ("sr", "<i8"), ("ss", "<i8"),
                           ("begin", "<i8"), ("O", "<i4"), ("I", "<i4"), ("KH", "<i4"), ("KW", "<i4"), ("opad", "<i4"),
                           ("ipad", "<i4"), ("krow", "<i4"), ("flip", "<i4"), ("row_scale", "<u8")])
            assert dt.itemsize == 104
            arr = np.zeros(len(jobs), dtype=dt)
            begin = 0
            for n, (src, out, O, I, KH, KW, opad, ipad, krow, so, si, sr, ss, flip) in enumerate(jobs):
                arr[n] = (src.data_ptr(), out.hi_ptr, out.lo_ptr or 0, so, si, sr, ss, begin, O, I, KH, KW, opad, ipad, krow, flip, 0)
                begin += -(-opad // 16) * -(-ipad // 64) * -(-(KH * KW) // 9)  # tiles of 16 rows x 64 channels x 9 taps (pack.cu)
            tab = dict(jobs=torch.from_numpy(arr.view(np.uint8).copy()).to(self.dev), njobs=len(jobs), blocks=begin,
                       keep=[j[0] for j in jobs] + [j[1] for j in jobs])  # sources and destinations stay alive with the table
            tables[key] = tab
        with torch.no_grad():
            torch.cat([self.model.pnp_net.fc_r.weight, self.model.pnp_net.fc_t.weight], 0, out=self.w_rt)
            torch.cat([self.model.pnp_net.fc_r.bias, self.model.pnp_net.fc_t.bias], 0, out=self.b_rt)
        C.gdrn_pack_weight_batched(tab["jobs"].data_ptr(), tab["njobs"], tab["blocks"], _stream())

    # ------------------------------------------------------------------------------------------ eval: folded conv+BN+ReLU
    # (conv key, BatchNorm key) of every conv that is followed by a BatchNorm, in forward order
    def _conv_bn_pairs(self):
        m = self.model
        out = [("stem", "backbone.bn1")]
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(m.backbone, f"layer{li}")):
                p = f"backbone.layer{li}.{bi}"
                out += [(p + ".conv1", p + ".bn1"), (p + ".conv2", p + ".bn2")]
                if blk.downsample is not None:
                    out.append((p + ".downsample.0", p + ".downsample.1"))
        out.append(("deconv", "rot_head_net.features.1"))
        out += [(f"rot_head_net.features.{ci}", f"rot_head_net.features.{bi}") for ci, bi, _u in HEAD_CONVS]
        return out

    def prepare_weights_folded(self):
        """Inference operands: eval-mode BatchNorm folded into the conv weights (scale, at pack time) and the conv epilogue
        (shift as bias), so conv + BN (+ identity) + ReLU is ONE kernel per layer.  Two launches per forward:
        gdrn_bn_fold_batched (all layers' scale / shift from the running statistics) and one batched weight pack."""
        import numpy as np

        key = tuple(p.data_ptr() for _, p in self.named_params)
        fold = getattr(self, "_fold", None)
        if fold is None or fold["key"] != key:
            pl = self.planes
            pairs = self._conv_bn_pairs()
            convs = dict(self._conv_modules())
            wfe, sc, sh = {}, {}, {}
            fdt = np.dtype([("gamma", "<u8"), ("beta", "<u8"), ("mean", "<u8"), ("var", "<u8"), ("scale", "<u8"), ("shift", "<u8"),
                            ("C", "<i4"), ("eps", "<f4"), ("pad", "<i8")])
            assert fdt.itemsize == 64
            farr = np.zeros(len(pairs), dtype=fdt)
            for n, (ck, bk) in enumerate(pairs):
                mod = self._bn_mods[bk]
                sc[bk] = torch.empty(mod.num_features, device=self.dev)
                sh[bk] = torch.empty(mod.num_features, device=self.dev)
                farr[n] = (mod.weight.data_ptr(), mod.bias.data_ptr(), mod.running_mean.data_ptr(), mod.running_var.data_ptr(),
                           sc[bk].data_ptr(), sh[bk].data_ptr(), mod.num_features, float(mod.eps), 0)
            ops._pack_recorder = []
            try:
                scales = []
                for ck, bk in pairs:
                    if ck == "stem":
                        w = self.model.backbone.conv1.weight
                        wfe[ck] = PT((64, 192), pl, device=self.dev, zero=True)
                        ops._pack(w, wfe[ck], 64, 3, 7, 7, 64, 3, 192, 147, 49, 7, 1, 0)
                    elif ck == "deconv":
                        pk = ops.pack_deconv_fwd_phases if _DECONV_PHASES else ops.pack_deconv_fwd
                        wfe[ck] = pk(self.model.rot_head_net.features[0].weight, pl)
                    else:
                        wfe[ck] = ops.pack_conv_fwd(convs[ck].weight, pl)
                    scales.append(sc[bk])
                jobs = ops._pack_recorder
            finally:
                ops._pack_recorder = None
            dt = np.dtype([("src", "<u8"), ("hi", "<u8"), ("lo", "<u8"), ("so", "<i8"), ("si", "<i8"), ("sr", "<i8"), ("ss", "<i8"),
                           ("begin", "<i8"), ("O", "<i4"), ("I", "<i4"), ("KH", "<i4"), ("KW", "<i4"), ("opad", "<i4"),
                           ("ipad", "<i4"), ("krow", "<i4"), ("flip", "<i4"), ("row_scale", "<u8")])
            assert dt.itemsize == 104
            arr = np.zeros(len(jobs), dtype=dt)
            begin = 0
            for n, (src, out, O, I, KH, KW, opad, ipad, krow, so, si, sr, ss, flip) in enumerate(jobs):
                arr[n] = (src.data_ptr(), out.hi_ptr, out.lo_ptr or 0, so, si, sr, ss, begin, O, I, KH, KW, opad, ipad, krow, flip,
                          scales[n].data_ptr())
                begin += -(-opad // 16) * -(-ipad // 64) * -(-(KH * KW) // 9)
            fold = self._fold = dict(key=key, wf=wfe, scale=sc, shift=sh, njobs=len(jobs), blocks=begin, nfold=len(pairs),
                                     fjobs=torch.from_numpy(farr.view(np.uint8).copy()).to(self.dev),
                                     jobs=torch.from_numpy(arr.view(np.uint8).copy()).to(self.dev), keep=[j[0] for j in jobs])
        C.gdrn_bn_fold_batched(fold["fjobs"].data_ptr(), fold["nfold"], _stream())
        C.gdrn_pack_weight_batched(fold["jobs"].data_ptr(), fold["njobs"], fold["blocks"], _stream())
        return fold

    def forward_eval_folded(self, x: torch.Tensor, aux: dict, want_maps: bool = False) -> dict:
        """Inference forward (reference caller gdrn_evaluator.py:568-580) with every conv + BatchNorm (+ residual) + ReLU
        fused into the GEMM epilogue: no BatchNorm pass, no pre-activation tensor in HBM."""
        m, pl, dev = self.model, self.planes, self.dev
        B = x.shape[0]
        assert x.shape[1:] == (3, 256, 256), x.shape
        fold_box = {}

        def packs():  # weight operands are input-independent: packed on the side stream while the main stream builds the im2col matrix
            self.prepare_weights(need_dgrad=False)  # head 1x1 / Patch-PnP / FC operands (no BatchNorm behind them)
            fold_box["F"] = self.prepare_weights_folded()

        self._on_side(packs)

        def cbr(xin, ck, bk, conv, relu=True, res=None, kind="conv"):
            if kind == "deconv":
                if _DECONV_PHASES:
                    return ops.conv_dgrad_s2(xin, wfe[ck], conv.out_channels, 3, 1, bias=shift[bk], act=2 if relu else 0)
                return ops.conv_fwd(ops.zero_insert(xin), wfe[ck], conv.out_channels, 3, 3, 1, 1, bias=shift[bk], act=2 if relu else 0,
                                    res=res, algo_scale=0.25)
            k, stride, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
            return ops.conv_fwd(xin, wfe[ck], conv.out_channels, k, k, stride, pad, bias=shift[bk], act=2 if relu else 0, res=res)

        a_col = PT((B * 128 * 128, 192), pl, device=dev)
        C.gdrn_stem_im2col(x.data_ptr(), a_col.hi_ptr, a_col.lo_ptr, B, 256, 256, _stream())
        self._join_side(force=True)
        wfe, shift = fold_box["F"]["wf"], fold_box["F"]["shift"]
        a0 = ops.gemm_fwd(a_col, wfe["stem"], 64, bias=shift["backbone.bn1"], act=2).view(B, 128, 128, 64)
        cur = ops.maxpool_fwd(a0)
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(m.backbone, f"layer{li}")):
                p = f"backbone.layer{li}.{bi}"
                a1 = cbr(cur, p + ".conv1", p + ".bn1", blk.conv1)
                idn = cur
                if blk.downsample is not None:
                    idn = cbr(cur, p + ".downsample.0", p + ".downsample.1", blk.downsample[0], relu=False)
                cur = cbr(a1, p + ".conv2", p + ".bn2", blk.conv2, res=idn)
        hf = m.rot_head_net.features
        cur = cbr(cur, "deconv", "rot_head_net.features.1", hf[0], kind="deconv")
        for ci, bi, up in HEAD_CONVS:
            if up:
                cur = ops.upsample2x_fwd(cur)
            cur = cbr(cur, f"rot_head_net.features.{ci}", f"rot_head_net.features.{bi}", hf[ci])
        return self._tail_inference(cur, aux, want_maps)

    def _eval_graphed(self, x: torch.Tensor, aux: dict, want_maps: bool) -> dict:
        """Inference forward replayed as ONE CUDA graph per input signature (`use_cuda_graphs`): ~230 launches incl. the
        BatchNorm folding and the weight pack (so updated weights / running statistics are picked up at every replay)."""
        tens = {k: v for k, v in aux.items() if isinstance(v, torch.Tensor)}
        key = (tuple(x.shape), want_maps, tuple(sorted((k, tuple(v.shape)) for k, v in tens.items())),
               tuple(p.data_ptr() for _, p in self.named_params))
        cache = self.__dict__.setdefault("_eval_graphs", {})
        g = cache.get(key)
        if g is None:
            st = dict(x=x.clone(), aux={k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in aux.items()})
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.forward_eval_folded(st["x"], st["aux"], want_maps=want_maps)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            st["graph"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st["graph"], capture_error_mode=_capture_mode()):
                st["res"] = self.forward_eval_folded(st["x"], st["aux"], want_maps=want_maps)
            if len(cache) > 8:
                cache.clear()
            g = cache[key] = st
        g["x"].copy_(x, non_blocking=True)
        for k, v in tens.items():
            g["aux"][k].copy_(v, non_blocking=True)
        g["graph"].replay()
        return {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in g["res"].items()}

    def _tail_inference(self, head_in: PT, aux: dict, want_maps: bool) -> dict:
        """1x1 output conv -> glue -> Patch-PnP -> test-time pose decode (shared by the folded and the plain eval forward)."""
        m, pl, dev = self.model, self.planes, self.dev
        B = head_in.shape[0]
        hf, pn, pf = m.rot_head_net.features, m.pnp_net, m.pnp_net.features
        logits = torch.empty(B * 4096, 72, device=dev)
        ops.conv_fwd(head_in, self.wf["rot_head_net.features.23"], 69, 1, 1, 1, 0, out_f32=logits, bias=hf[23].bias, ldc=72,
                     want_planes=False)
        pnp_in = PT((B, 64, 64, 128), pl, device=dev)
        C.gdrn_head_glue_fwd(logits.data_ptr(), ops.ptr(aux.get("roi_coord_2d")), aux["roi_extents"].data_ptr(), pnp_in.hi_ptr,
                             pnp_in.lo_ptr, B, 4096, self.with_2d, _stream())
        cur = pnp_in
        for ci, gi in ((0, 1), (3, 4), (6, 7)):
            u = ops.conv_fwd(cur, self.wf[f"pnp_net.features.{ci}"], 128, 3, 3, 2, 1)
            gstats = torch.empty(B, 32, 2, device=dev)
            cur = ops.gn_relu_fwd(u, pf[gi].weight, pf[gi].bias, gstats, G=pf[gi].num_groups, eps=pf[gi].eps)
        h1 = ops.gemm_fwd(cur.view(B, 8192), self.wf["fc1"], 1024, bias=pn.fc1.bias, act=1)
        h2 = ops.gemm_fwd(h1, self.wf["fc2"], 256, bias=pn.fc2.bias, act=1)
        pred = torch.zeros(B, 16, device=dev)
        ops.gemm_fwd(h2, self.wf["fc_rt"], 9, out_f32=pred, bias=self.b_rt, ldc=16, want_planes=False)
        out_rot = torch.empty(B, 3, 3, device=dev)
        out_trans = torch.empty(B, 3, device=dev)
        C.gdrn_pose_loss(pred.data_ptr(), 16, aux["roi_cams"].data_ptr(), aux["roi_centers"].data_ptr(),
                         aux["roi_whs"].data_ptr(), aux["resize_ratios"].data_ptr(), aux["roi_extents"].data_ptr(), None,
                         None, None, None, None, None, None, out_rot.data_ptr(), out_trans.data_ptr(), None, None, None,
                         None, B, 0, 0, 0.0, _stream())
        res = dict(rot=out_rot, trans=out_trans, pred=pred, logits=logits)
        if want_maps:
            maps = logits.view(B, 64, 64, 72).permute(0, 3, 1, 2)
            res.update(mask=maps[:, 0:1], coor_x=maps[:, 1:2], coor_y=maps[:, 2:3], coor_z=maps[:, 3:4], region=maps[:, 4:69])
        return res

    def _prepare_weights_calls(self, need_dgrad: bool):
        m, pl, bpl = self.model, self.planes, self.bplanes
        wf, wd = self.wf, self.wd
        for key, conv in self._conv_modules():
            if key == "backbone.conv1":
                continue  # the 7x7 stem is a GEMM over its im2col matrix (_pack_stem)
            wf[key] = ops.pack_conv_fwd(conv.weight, pl, out=wf.get(key))
            if need_dgrad:
                wd[key] = ops.pack_conv_dgrad(conv.weight, bpl, out=wd.get(key))
        dc = m.rot_head_net.features[0]
        wf["deconv"] = (ops.pack_deconv_fwd_phases if _DECONV_PHASES else ops.pack_deconv_fwd)(dc.weight, pl, out=wf.get("deconv"))
        if need_dgrad:
            wd["deconv"] = ops.pack_deconv_dgrad(dc.weight, bpl, out=wd.get("deconv"))
        pn = m.pnp_net
        wf["fc1"] = ops.pack_linear(pn.fc1.weight, pl, out=wf.get("fc1"), nhwc_from=(128, 8, 8))
        wf["fc2"] = ops.pack_linear(pn.fc2.weight, pl, out=wf.get("fc2"))
        if not hasattr(self, "w_rt"):
            self.w_rt = torch.empty(pn.fc_r.out_features + 3, pn.fc_r.in_features, device=self.dev)
            self.b_rt = torch.empty(pn.fc_r.out_features + 3, device=self.dev)
        wf["fc_rt"] = ops.pack_linear(self.w_rt, pl, out=wf.get("fc_rt"))
        if need_dgrad:
            wd["fc1"] = ops.pack_linear(pn.fc1.weight, bpl, out=wd.get("fc1"), nhwc_from=(128, 8, 8), transpose=True)
            wd["fc2"] = ops.pack_linear(pn.fc2.weight, bpl, out=wd.get("fc2"), transpose=True)
            wd["fc_rt"] = ops.pack_linear(self.w_rt, bpl, out=wd.get("fc_rt"), transpose=True)

    def _pack_stem(self):
        """7x7 stem weight [64][3][7][7] -> [64][192] with k = (r*7+s)*3 + c (matches gdrn_stem_im2col)."""
        w = self.model.backbone.conv1.weight
        out = self.wf.get("stem") or PT((64, 192), self.planes, device=self.dev, zero=True)  # cols 147..191 stay zero
        ops._pack(w, out, 64, 3, 7, 7, 64, 3, 192, 147, 49, 7, 1, 0)
        self.wf["stem"] = out

    # ------------------------------------------------------------------------------------------ building blocks
    def _bn_fwd(self, key: str, u: PT, relu: bool, train_bn: bool, res: Optional[PT] = None) -> PT:
        """y = [relu](bn(u) [+ res]).  When the step keeps activations for backward (`self._want_masks`) and relu is set, the
        kernel also emits the ReLU mask as one BIT per element (self._last_mask): BatchNorm backward then reads 1/8 byte
        instead of the 2-byte activation per element in both of its passes."""
        mod, st = self._bn_mods[key], self.bn[key]
        Cc = mod.num_features
        count = u.numel() // Cc
        mom = mod.momentum if mod.momentum is not None else 0.1
        if train_bn and mod.num_batches_tracked is not None:
            self._nbt.append(mod.num_batches_tracked)
        if train_bn and self.deterministic:
            C.gdrn_bn_stats(u.hi_ptr, u.lo_ptr, self._det_workspace().data_ptr(), st.stats.data_ptr(), count, Cc, _stream())
        y = ops.like(u)
        mask = None
        if relu and getattr(self, "_want_masks", False) and _BN_BITMASK:
            mask = torch.empty(u.numel() // 8, dtype=torch.uint8, device=self.dev)
        self._last_mask = mask
        C.gdrn_bn_fwd(u.hi_ptr, u.lo_ptr, res.hi_ptr if res else None, res.lo_ptr if res else None, y.hi_ptr, y.lo_ptr,
                      ops.ptr(st.stats) if train_bn else None, mod.weight.data_ptr(), mod.bias.data_ptr(),
                      mod.running_mean.data_ptr(), mod.running_var.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(),
                      ops.ptr(mask), count, Cc, float(mod.eps), float(mom), int(train_bn), int(relu), _stream())
        return y

    def _det_workspace(self):
        if self._det_ws is None:
            self._det_ws = torch.empty(2 * 512 * int(C.load().gdrn_det_parts()), device=self.dev)
        return self._det_ws

    def _conv_bn(self, x: PT, ckey: str, conv, bkey: str, relu: bool, train_bn: bool, res: Optional[PT] = None,
                 wkey: Optional[str] = None, kind: str = "conv"):
        st = self.bn[bkey]
        stats = st.stats if (train_bn and not self.deterministic) else None  # deterministic: statistics from gdrn_bn_stats
        k, stride, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        if kind == "deconv" and _DECONV_PHASES:
            u = ops.conv_dgrad_s2(x, self.wf["deconv"], conv.out_channels, 3, 1, stats=stats)
        elif kind == "deconv":
            u = ops.conv_fwd(x, self.wf["deconv"], conv.out_channels, 3, 3, 1, 1, stats=stats, algo_scale=0.25)
        else:
            u = ops.conv_fwd(x, self.wf[wkey or ckey], conv.out_channels, k, k, stride, pad, stats=stats)
        y = self._bn_fwd(bkey, u, relu, train_bn, res)
        return u, y

    # ------------------------------------------------------------------------------------------ forward
    def run(self, x, *, roi_coord_2d, roi_cams, roi_centers, roi_whs, roi_extents, resize_ratios, gt_xyz=None,
            gt_mask_trunc=None, gt_mask_visib=None, gt_region=None, gt_ego_rot=None, gt_points=None, sym_infos=None,
            gt_trans=None, gt_trans_ratio=None, do_loss=False, train_bn=False, want_maps=False):
        aux = dict(roi_coord_2d=roi_coord_2d, roi_cams=roi_cams, roi_centers=roi_centers, roi_whs=roi_whs,
                   roi_extents=roi_extents, resize_ratios=resize_ratios, gt_xyz=gt_xyz, gt_mask_trunc=gt_mask_trunc,
                   gt_mask_visib=gt_mask_visib, gt_region=gt_region, gt_ego_rot=gt_ego_rot, gt_points=gt_points,
                   sym_infos=sym_infos, gt_trans=gt_trans, gt_trans_ratio=gt_trans_ratio)
        B = x.shape[0]
        for k_, v in aux.items():
            if isinstance(v, torch.Tensor):
                v = v.to(self.dev)
                # gt_region is handed to the kernels as `const long long*` (the reference calls .long() itself, GDRN.py:393,
                # and its dataset emits int32): convert ANY label dtype explicitly; everything else is fp32
                aux[k_] = (v.long() if k_ == "gt_region" else v.float()).contiguous()
        if roi_cams is not None and aux["roi_cams"].dim() == 2:  # one shared K: broadcast like the reference does
            aux["roi_cams"] = aux["roi_cams"].unsqueeze(0).expand(B, 3, 3).contiguous()
        want = dict(roi_coord_2d=(B, 2, 64, 64), roi_cams=(B, 3, 3), roi_centers=(B, 2), roi_whs=(B, 2), roi_extents=(B, 3),
                    resize_ratios=(B,), gt_xyz=(B, 3, 64, 64), gt_mask_trunc=(B, 64, 64), gt_mask_visib=(B, 64, 64),
                    gt_region=(B, 64, 64), gt_ego_rot=(B, 3, 3), gt_trans=(B, 3), gt_trans_ratio=(B, 3))
        for k_, shp in want.items():
            v = aux.get(k_)
            if isinstance(v, torch.Tensor):
                if v.numel() != int(torch.Size(shp).numel()):
                    raise ValueError(f"{k_}: expected shape {shp} for a batch of {B}, got {tuple(v.shape)}")
                aux[k_] = v.view(shp)
        if isinstance(aux.get("gt_points"), torch.Tensor) and (aux["gt_points"].dim() != 3 or aux["gt_points"].shape[0] != B
                                                               or aux["gt_points"].shape[2] != 3):
            raise ValueError(f"gt_points: expected [B={B}, n, 3], got {tuple(aux['gt_points'].shape)}")
        if aux.get("sym_infos") is not None:
            if len(aux["sym_infos"]) != B:
                raise ValueError(f"sym_infos: expected a list of {B} entries, got {len(aux['sym_infos'])}")
            aux["sym_idx"] = self.sym_table.indices(aux["sym_infos"])  # the matrices themselves are already on the device
        aux.pop("sym_infos", None)
        x = x.to(self.dev).float().contiguous()
        if not do_loss:
            with torch.no_grad():
                if not train_bn and self.fold_eval:  # inference: conv + BN (+ identity) + ReLU fused per layer
                    if self.use_cuda_graphs:
                        return self._eval_graphed(x, aux, want_maps)
                    return self.forward_eval_folded(x, aux, want_maps=want_maps)
                return self.forward(x, aux, train_bn=train_bn, do_loss=False, want_maps=want_maps)
        params = [p for _, p in self.named_params]
        losses, vis = _GDRNFunction.apply(self, x, aux, train_bn, *params)
        return dict(losses=losses, vis=vis)

    def forward(self, x: torch.Tensor, aux: dict, train_bn: bool, do_loss: bool, want_maps: bool = False) -> dict:
        m, pl, dev = self.model, self.planes, self.dev
        B = x.shape[0]
        assert x.shape[1:] == (3, 256, 256), x.shape
        # the per-step weight re-pack (0.24 ms) does not depend on the input: it runs on the side stream while the main
        # stream builds the stem's im2col matrix, and is joined right before the first GEMM
        self._on_side(lambda: self.prepare_weights(need_dgrad=do_loss))
        self._nbt = []
        if train_bn:
            self.stats_all.zero_()
        S = dict(B=B, train_bn=train_bn) if do_loss else None
        self._want_masks = do_loss

        # ---- a1: backbone (resnet_backbone.py:69-76)
        a_col = PT((B * 128 * 128, 192), pl, device=dev)
        C.gdrn_stem_im2col(x.data_ptr(), a_col.hi_ptr, a_col.lo_ptr, B, 256, 256, _stream())
        self._join_side(force=True)
        st0 = self.bn["backbone.bn1"]
        u0 = ops.gemm_fwd(a_col, self.wf["stem"], 64, stats=st0.stats if (train_bn and not self.deterministic) else None).view(B, 128, 128, 64)
        a0 = self._bn_fwd("backbone.bn1", u0, True, train_bn)
        if S is not None:
            cur, pool_arg = ops.maxpool_fwd(a0, want_arg=True)
            S["stem"] = dict(a_col=a_col, u0=u0, a0=a0, pool_arg=pool_arg, m0=self._last_mask)
            S["blocks"] = []
        else:
            cur = ops.maxpool_fwd(a0)
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(m.backbone, f"layer{li}")):
                p = f"backbone.layer{li}.{bi}"
                u1, a1 = self._conv_bn(cur, p + ".conv1", blk.conv1, p + ".bn1", True, train_bn)
                m1 = self._last_mask
                ud = None
                if blk.downsample is not None:
                    ud, idn = self._conv_bn(cur, p + ".downsample.0", blk.downsample[0], p + ".downsample.1", False, train_bn)
                else:
                    idn = cur
                u2, out = self._conv_bn(a1, p + ".conv2", blk.conv2, p + ".bn2", True, train_bn, res=idn)
                if S is not None:
                    S["blocks"].append(dict(p=p, blk=blk, x_in=cur, u1=u1, a1=a1, u2=u2, out=out, ud=ud, m1=m1, m2=self._last_mask))
                cur = out
        feat = cur  # [B,8,8,512]

        # ---- a2: geometry head (cdpn_rot_head_region.py:80-135)
        hf = m.rot_head_net.features
        z = feat if _DECONV_PHASES else ops.zero_insert(feat)
        u, y = self._conv_bn(z, "deconv", hf[0], "rot_head_net.features.1", True, train_bn, kind="deconv")
        if S is not None:
            S["deconv"] = dict(z=z, u=u, y=y, m=self._last_mask)
            S["head"] = []
        cur = y
        for ci, bi, up in HEAD_CONVS:
            if up:
                cur = ops.upsample2x_fwd(cur)
            u, y = self._conv_bn(cur, f"rot_head_net.features.{ci}", hf[ci], f"rot_head_net.features.{bi}", True, train_bn)
            if S is not None:
                S["head"].append(dict(ci=ci, bi=bi, up=up, x_in=cur, u=u, y=y, m=self._last_mask))
            cur = y
        head_in = cur  # [B,64,64,256]
        if self._nbt:
            torch._foreach_add_(self._nbt, 1)  # all BatchNorm num_batches_tracked counters in one launch
        logits = torch.empty(B * 4096, 72, device=dev)
        ops.conv_fwd(head_in, self.wf["rot_head_net.features.23"], 69, 1, 1, 1, 0, out_f32=logits, bias=hf[23].bias, ldc=72,
                     want_planes=False)

        # ---- a3 + a4: glue and Patch-PnP (GDRN.py:156-181, conv_pnp_net.py:111-157)
        pnp_in = PT((B, 64, 64, 128), pl, device=dev)
        C.gdrn_head_glue_fwd(logits.data_ptr(), ops.ptr(aux.get("roi_coord_2d")), aux["roi_extents"].data_ptr(), pnp_in.hi_ptr,
                             pnp_in.lo_ptr, B, 4096, self.with_2d, _stream())
        pf = m.pnp_net.features
        cur = pnp_in
        pnp_saved = []
        for ci, gi in ((0, 1), (3, 4), (6, 7)):
            u = ops.conv_fwd(cur, self.wf[f"pnp_net.features.{ci}"], 128, 3, 3, 2, 1)
            gstats = torch.empty(B, 32, 2, device=dev)
            y = ops.gn_relu_fwd(u, pf[gi].weight, pf[gi].bias, gstats, G=pf[gi].num_groups, eps=pf[gi].eps)
            pnp_saved.append(dict(ci=ci, gi=gi, x_in=cur, u=u, y=y, gstats=gstats))
            cur = y
        flat = cur.view(B, 8192)
        pn = m.pnp_net
        h1 = ops.gemm_fwd(flat, self.wf["fc1"], 1024, bias=pn.fc1.bias, act=1)
        h2 = ops.gemm_fwd(h1, self.wf["fc2"], 256, bias=pn.fc2.bias, act=1)
        pred = torch.zeros(B, 16, device=dev)  # cols 0..5 rot6d, 6..8 (centroid dx, dy, z)
        ops.gemm_fwd(h2, self.wf["fc_rt"], 9, out_f32=pred, bias=self.b_rt, ldc=16, want_planes=False)

        # ---- a5/a6 pose decode (+ a7..a10 losses)
        out_rot = torch.empty(B, 3, 3, device=dev)
        out_trans = torch.empty(B, 3, device=dev)
        res = dict(rot=out_rot, trans=out_trans, pred=pred)
        if not do_loss:
            C.gdrn_pose_loss(pred.data_ptr(), 16, aux["roi_cams"].data_ptr(), aux["roi_centers"].data_ptr(),
                             aux["roi_whs"].data_ptr(), aux["resize_ratios"].data_ptr(), aux["roi_extents"].data_ptr(), None,
                             None, None, None, None, None, None, out_rot.data_ptr(), out_trans.data_ptr(), None, None, None,
                             None, B, 0, 0, 0.0, _stream())
            if want_maps:
                maps = logits.view(B, 64, 64, 72).permute(0, 3, 1, 2)
                res.update(mask=maps[:, 0:1], coor_x=maps[:, 1:2], coor_y=maps[:, 2:3], coor_z=maps[:, 3:4], region=maps[:, 4:69])
            res["logits"] = logits
            return res

        pix_sums = torch.empty(6, dtype=torch.float64, device=dev)
        C.gdrn_pixel_loss_fwd(logits.data_ptr(), aux["gt_xyz"].data_ptr(), aux["gt_mask_visib"].data_ptr(),
                              aux["gt_mask_trunc"].data_ptr(), aux["gt_region"].data_ptr(), pix_sums.data_ptr(), B, 4096, _stream())
        S.update(logits=logits, pnp_in=pnp_in, pnp=pnp_saved, flat=flat, h1=h1, h2=h2, pred=pred, head_in=head_in,
                 pix_sums=pix_sums, aux=aux)
        self.saved = S
        # the pose kernel needs the upstream loss gradients to emit dY9; forward calls it with unit weights and
        # backward re-runs it (cheap: B CTAs) with the actual ones.
        self._pose(S, gw=None)
        losses = torch.empty(8, device=dev)
        vis2 = torch.empty(2, device=dev)
        n_pts = aux["gt_points"].shape[1]
        C.gdrn_loss_finalize(pix_sums.data_ptr(), S["pose_sums"].data_ptr(), S["vis_ps"].data_ptr(), losses.data_ptr(),
                             vis2.data_ptr(), B, 4096, n_pts, _stream())
        vis = torch.cat([vis2, out_trans_first(S["out_trans"]), pred[0, 6:9], aux["gt_trans"][0], aux["gt_trans_ratio"][0]])
        res.update(rot=S["out_rot"], trans=S["out_trans"], losses=losses, vis=vis, logits=logits)
        return res

    def _pose(self, S, gw):
        aux, B, dev = S["aux"], S["B"], self.dev
        if "out_rot" not in S:
            S["out_rot"] = torch.empty(B, 3, 3, device=dev)
            S["out_trans"] = torch.empty(B, 3, device=dev)
            S["pose_sums"] = torch.empty(4, dtype=torch.float64, device=dev)
            S["vis_ps"] = torch.empty(B, 2, device=dev)
            S["dy9"] = PT((B, 64), self.bplanes, device=dev)
        sym_idx = aux.get("sym_idx")
        if sym_idx is None and aux.get("sym_infos") is not None:  # direct engine.forward callers may pass the raw list
            sym_idx = aux["sym_idx"] = self.sym_table.indices(aux["sym_infos"])
        syms = self.sym_table.table if sym_idx is not None else None
        if gw is None:
            gw = torch.ones(3, device=dev)
        n_pts = aux["gt_points"].shape[1]
        C.gdrn_pose_loss(S["pred"].data_ptr(), 16, aux["roi_cams"].data_ptr(), aux["roi_centers"].data_ptr(),
                         aux["roi_whs"].data_ptr(), aux["resize_ratios"].data_ptr(), aux["roi_extents"].data_ptr(),
                         aux["gt_points"].data_ptr(), aux["gt_ego_rot"].data_ptr(), aux["gt_trans"].data_ptr(),
                         aux["gt_trans_ratio"].data_ptr(), ops.ptr(syms), ops.ptr(sym_idx), gw.data_ptr(),
                         S["out_rot"].data_ptr(), S["out_trans"].data_ptr(), S["pose_sums"].data_ptr(), S["vis_ps"].data_ptr(),
                         S["dy9"].hi_ptr, S["dy9"].lo_ptr, B, n_pts, 1, 1e-4, _stream())

    # ------------------------------------------------------------------------------------------ backward
    # ---- weight gradients on a SIDE STREAM.  In backward the chain  BN-bwd(L) -> dgrad(L) -> BN-bwd(L-1) -> ...  alternates
    # tensor-bound GEMMs with HBM-bound BatchNorm passes, while wgrad(L) (+ its split-K unpack) only needs du(L) and feeds
    # nothing but the flat gradient buffer.  Issued on a second stream it runs concurrently with dgrad(L) and -- the point --
    # with the HBM-bound BN-bwd(L-1): tensor-bound and bandwidth-bound work overlap instead of serialising.  The streams join
    # at every gradient-bucket boundary (before the loss-scale removal / all-reduce of that bucket) and at the end of backward;
    # inside a CUDA graph the fork / join become graph edges.  GDRN_WGRAD_STREAM=0 keeps everything on one stream (A/B).
    def _on_side(self, fn, keep=()):
        if not self.wgrad_side_stream:
            return fn()
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        side = self._side_stream
        side.wait_stream(torch.cuda.current_stream())  # du(L) is ready
        with torch.cuda.stream(side):
            fn()
        self._side_keep.extend(keep)  # operands allocated on the main stream stay referenced until the join

    def _join_side(self, force: bool = False):
        if self._side_stream is not None and (self._side_keep or force):
            torch.cuda.current_stream().wait_stream(self._side_stream)
        self._side_keep.clear()

    def _wgrad_conv(self, du: PT, x_in: PT, conv, wname: str, ipad: Optional[int] = None):
        O, I, k = conv.out_channels, conv.in_channels, conv.kernel_size[0]

        def work():
            buf, ks, kss = ops.conv_wgrad(du, x_in, self.ws, O, k, k, conv.stride[0], conv.padding[0])
            ops.unpack_wgrad(buf, self.grads[wname], O, I, k, k, ipad or x_in.shape[-1], ks, kss, I * k * k, k * k, k, 1)

        self._on_side(work, keep=(du, x_in))

    def _dgrad_conv(self, du: PT, conv, wkey: str) -> PT:
        k, stride, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        if stride == 2 and not _DGRAD_ZERO_INSERT:
            # four output-parity phases over the un-dilated dY (no zero-inserted tensor, 1/4 of its MACs)
            return ops.conv_dgrad_s2(du, self.wd[wkey], conv.in_channels, k, pad)
        z = ops.zero_insert(du) if stride == 2 else du
        return ops.conv_fwd(z, self.wd[wkey], conv.in_channels, k, k, 1, k - 1 - pad, algo_scale=0.25 if stride == 2 else 1.0)

    def _bn_bwd(self, bkey: str, ga: PT, gb: Optional[PT], y: Optional[PT], u: PT, want_gout=False, relu_from_u=False, mask=None):
        """mask: the forward's ReLU bitmap (preferred); else the mask comes from the activation y, or (relu_from_u, A/B
        switch) is recomputed from u for conv-BN-ReLU without a residual."""
        mod, st = self._bn_mods[bkey], self.bn[bkey]
        relu_from_u = relu_from_u and _BN_MASK_FROM_U and mask is None
        return ops.bn_bwd(ga, gb, None if (relu_from_u or mask is not None) else y, u, st.mean, st.invstd, mod.weight, st.sums,
                          self.grads[bkey + ".weight"], self.grads[bkey + ".bias"], self.saved["train_bn"], want_gout=want_gout,
                          beta=mod.bias, relu_from_u=relu_from_u, sums_zeroed=True, relu_mask=mask,
                          det_ws=self._det_workspace() if self.deterministic else None)

    def backward(self, grad_losses: torch.Tensor):
        """grad_losses: [8] upstream gradients of LOSS_NAMES.  Fills self.grads (views of self.flat_grad)."""
        S, m, dev = self.saved, self.model, self.dev
        assert S is not None, "backward called without a do_loss forward"
        B, aux, pn = S["B"], S["aux"], m.pnp_net
        bp = self.bplanes

        def h(t: PT) -> PT:  # a saved forward tensor as a backward operand
            return t.as_planes(bp)

        self.flat_grad.zero_()
        self.bwd_sums_all.zero_()
        gw = grad_losses.float().contiguous()
        if self.grad_scale != 1.0:
            gw = gw * self.grad_scale  # the whole backward is linear in the loss gradients
        self._pose(S, gw=gw[5:8].contiguous())
        dy9 = S["dy9"]
        tmp = torch.empty(1024, device=dev)

        # ---- FC stack (conv_pnp_net.py:152-156)
        buf, ks, kss = ops.gemm_wgrad(dy9, h(S["h2"]), self.ws)
        ops.unpack_wgrad(buf, self.grads["pnp_net.fc_r.weight"], 6, 256, 1, 1, 256, ks, kss, 256, 1, 0, 0)
        ops.unpack_wgrad(buf[6 * 256:], self.grads["pnp_net.fc_t.weight"], 3, 256, 1, 1, 256, ks, kss, 256, 1, 0, 0)
        C.gdrn_colsum(dy9.hi_ptr, dy9.lo_ptr, tmp.data_ptr(), B, 64, _stream())
        self.grads["pnp_net.fc_r.bias"].copy_(tmp[0:6])
        self.grads["pnp_net.fc_t.bias"].copy_(tmp[6:9])
        dh2 = ops.gemm_fwd(dy9, self.wd["fc_rt"], 256)
        dz2 = self._leaky_bwd(dh2, h(S["h2"]))
        buf, ks, kss = ops.gemm_wgrad(dz2, h(S["h1"]), self.ws)
        ops.unpack_wgrad(buf, self.grads["pnp_net.fc2.weight"], 256, 1024, 1, 1, 1024, ks, kss, 1024, 1, 0, 0)
        C.gdrn_colsum(dz2.hi_ptr, dz2.lo_ptr, self.grads["pnp_net.fc2.bias"].data_ptr(), B, 256, _stream())
        dh1 = ops.gemm_fwd(dz2, self.wd["fc2"], 1024)
        dz1 = self._leaky_bwd(dh1, h(S["h1"]))
        buf, ks, kss = ops.gemm_wgrad(dz1, h(S["flat"]), self.ws)
        ops.unpack_wgrad(buf, self.grads["pnp_net.fc1.weight"], 1024, 128, 8, 8, 128, ks, kss, 8192, 64, 8, 1)
        C.gdrn_colsum(dz1.hi_ptr, dz1.lo_ptr, self.grads["pnp_net.fc1.bias"].data_ptr(), B, 1024, _stream())
        g = ops.gemm_fwd(dz1, self.wd["fc1"], 8192).view(B, 8, 8, 128)

        # ---- Patch-PnP convs (conv_pnp_net.py:76-80)
        pf = pn.features
        for L in reversed(S["pnp"]):
            ci, gi = L["ci"], L["gi"]
            du = ops.gn_relu_bwd(g, h(L["y"]), h(L["u"]), pf[gi].weight, L["gstats"], self.grads[f"pnp_net.features.{gi}.weight"],
                                 self.grads[f"pnp_net.features.{gi}.bias"], G=pf[gi].num_groups)
            self._wgrad_conv(du, h(L["x_in"]), pf[ci], f"pnp_net.features.{ci}.weight")
            g = self._dgrad_conv(du, pf[ci], f"pnp_net.features.{ci}")
        d_pnp_in = g  # [B,64,64,128]
        self._segment_done("pnp_net")

        # ---- glue + per-pixel losses backward (one fused pass over the logits)
        dlog = PT((B * 4096, 128), bp, device=dev)
        C.gdrn_head_bwd(S["logits"].data_ptr(), aux["gt_xyz"].data_ptr(), aux["gt_mask_visib"].data_ptr(),
                        aux["gt_mask_trunc"].data_ptr(), aux["gt_region"].data_ptr(), S["pix_sums"].data_ptr(), gw.data_ptr(),
                        d_pnp_in.hi_ptr, d_pnp_in.lo_ptr, aux["roi_extents"].data_ptr(), dlog.hi_ptr, dlog.lo_ptr, B, 4096,
                        self.with_2d, _stream())
        hf = m.rot_head_net.features
        head_in = h(S["head_in"])
        def head_out_wgrad():
            buf, ks, kss = ops.gemm_wgrad(dlog, head_in.view(B * 4096, 256), self.ws)
            ops.unpack_wgrad(buf, self.grads["rot_head_net.features.23.weight"], 69, 256, 1, 1, 256, ks, kss, 256, 1, 0, 0)

        self._on_side(head_out_wgrad, keep=(dlog,))
        tmp128 = torch.empty(128, device=dev)
        C.gdrn_colsum(dlog.hi_ptr, dlog.lo_ptr, tmp128.data_ptr(), B * 4096, 128, _stream())
        self.grads["rot_head_net.features.23.bias"].copy_(tmp128[:69])
        g = ops.gemm_fwd(dlog, self.wd["rot_head_net.features.23"], 256).view(B, 64, 64, 256)

        # ---- head convs (cdpn_rot_head_region.py:95-125)
        for L in reversed(S["head"]):
            ci, bi = L["ci"], L["bi"]
            du, _ = self._bn_bwd(f"rot_head_net.features.{bi}", g, None, h(L["y"]), h(L["u"]), relu_from_u=True, mask=L["m"])
            self._wgrad_conv(du, h(L["x_in"]), hf[ci], f"rot_head_net.features.{ci}.weight")
            g = self._dgrad_conv(du, hf[ci], f"rot_head_net.features.{ci}")
            if L["up"]:
                g = ops.upsample2x_bwd(g)
        D = S["deconv"]
        du, _ = self._bn_bwd("rot_head_net.features.1", g, None, h(D["y"]), h(D["u"]), relu_from_u=True, mask=D["m"])
        def deconv_wgrad(du=du):
            if _DECONV_PHASES:
                # d wt[ci][co][r][s] = sum_q x[ci][q] * du[co][2q + (r, s) - 1]: the weight gradient of the stride-2 conv du -> x
                # with x in the role of dY; its [O = 512][I = 256][3][3] layout IS the ConvTranspose2d IOHW layout (no flip)
                buf, ks, kss = ops.conv_wgrad(h(D["z"]), du, self.ws, 512, 3, 3, 2, 1)
                ops.unpack_wgrad(buf, self.grads["rot_head_net.features.0.weight"], 512, 256, 3, 3, 256, ks, kss, 256 * 9, 9, 3, 1)
                return
            buf, ks, kss = ops.conv_wgrad(du, h(D["z"]), self.ws, 256, 3, 3, 1, 1)
            # ConvTranspose2d weight is IOHW [512][256][3][3] with flipped taps relative to the equivalent conv
            ops.unpack_wgrad(buf, self.grads["rot_head_net.features.0.weight"], 256, 512, 3, 3, 512, ks, kss, 9, 256 * 9, 3, 1, flip=1)

        self._on_side(deconv_wgrad, keep=(du,))
        g = ops.conv_fwd(du, self.wd["deconv"], 512, 3, 3, 2, 1)  # [B,8,8,512]
        self._segment_done("rot_head_net")

        # ---- backbone (torchvision BasicBlock backward)
        ga, gb = g, None
        for Lb in reversed(S["blocks"]):
            p, blk = Lb["p"], Lb["blk"]
            du2, gout = self._bn_bwd(p + ".bn2", ga, gb, h(Lb["out"]), h(Lb["u2"]), want_gout=True, mask=Lb["m2"])
            self._wgrad_conv(du2, h(Lb["a1"]), blk.conv2, p + ".conv2.weight")
            da1 = self._dgrad_conv(du2, blk.conv2, p + ".conv2")
            du1, _ = self._bn_bwd(p + ".bn1", da1, None, h(Lb["a1"]), h(Lb["u1"]), relu_from_u=True, mask=Lb["m1"])
            self._wgrad_conv(du1, h(Lb["x_in"]), blk.conv1, p + ".conv1.weight")
            dx_main = self._dgrad_conv(du1, blk.conv1, p + ".conv1")
            if blk.downsample is not None:
                dud, _ = self._bn_bwd(p + ".downsample.1", gout, None, None, h(Lb["ud"]))
                self._wgrad_conv(dud, h(Lb["x_in"]), blk.downsample[0], p + ".downsample.0.weight")
                dx_ds = self._dgrad_conv(dud, blk.downsample[0], p + ".downsample.0")
                ga, gb = dx_main, dx_ds
            else:
                ga, gb = dx_main, gout
            if p.endswith(".0") and not p.startswith("backbone.layer1"):
                self._segment_done(p[:-2])  # all gradients of backbone.layerN are final
        g_pool = ops.add2(ga, gb)
        St = S["stem"]
        g_a0 = ops.maxpool_bwd(St["pool_arg"], g_pool)
        du0, _ = self._bn_bwd("backbone.bn1", g_a0, None, h(St["a0"]), h(St["u0"]), relu_from_u=True, mask=St["m0"])
        def stem_wgrad():  # on the side stream like every conv wgrad: the split-K workspace is owned by that stream by now
            buf, ks, kss = ops.gemm_wgrad(du0.view(B * 128 * 128, 64), h(St["a_col"]), self.ws)
            # ws rows [64][192] with k = (r*7+s)*3 + c  ->  OIHW [64][3][7][7]
            ops.unpack_wgrad(buf, self.grads["backbone.conv1.weight"], 64, 3, 7, 7, 3, ks, kss, 147, 49, 7, 1, krow=192)

        self._on_side(stem_wgrad, keep=(du0,))
        self._segment_done("backbone.stem")  # layer1 + bn1 + conv1 (joins the wgrad side stream)
        self.saved = None
        return self.grads

    def materialize_aliased_grads(self):
        """A p.grad that still aliases flat_grad (adopted by autograd in an earlier step and not reset by zero_grad) would be
        overwritten by the next backward: give such gradients their own storage first (gradient accumulation stays exact)."""
        lo = self.flat_grad.data_ptr()
        hi = lo + self.flat_grad.numel() * self.flat_grad.element_size()
        for _, p in self.named_params:
            g = p.grad
            if g is not None and lo <= g.data_ptr() < hi:
                p.grad = g.clone()

    def _segment_done(self, stage: str):
        self._join_side()
        self._segment_done_joined(stage)

    def _segment_done_joined(self, stage: str):
        """All gradients of a sub-network are final: remove the loss scale from its slice of the flat buffer, then hand
        it to the data-parallel hook (bucketed all-reduce on a side stream)."""
        if self.grad_scale != 1.0:
            names = ("backbone.conv1", "backbone.bn1", "backbone.layer1") if stage == "backbone.stem" else (stage,)
            lo = min(self._seg[n][0] for n in names)
            hi = max(self._seg[n][1] for n in names)
            C.gdrn_scale_f32(self.flat_grad.data_ptr() + 4 * lo, hi - lo, 1.0 / self.grad_scale, _stream())
        if self.grad_hook:
            self.grad_hook(self, stage)

    def _leaky_bwd(self, g: PT, y: PT) -> PT:
        out = ops.like(g)
        C.gdrn_leaky_bwd(g.hi_ptr, g.lo_ptr, y.hi_ptr, out.hi_ptr, out.lo_ptr, g.numel(), _stream())
        return out


def out_trans_first(t):
    return t[0]


def _bn_snapshot(engine: "Engine"):
    """BatchNorm running statistics + counters (graph warm-up iterations must not advance them)."""
    return [(m, m.running_mean.clone(), m.running_var.clone(),
             None if m.num_batches_tracked is None else m.num_batches_tracked.clone()) for m in engine._bn_mods.values()]


def _bn_restore(snap):
    with torch.no_grad():
        for m, rm, rv, nbt in snap:
            m.running_mean.copy_(rm)
            m.running_var.copy_(rv)
            if nbt is not None:
                m.num_batches_tracked.copy_(nbt)


def _capture_mode():
    # with a process group alive, NCCL's watchdog thread polls CUDA events while we capture: "thread_local" keeps its
    # calls legal (the collectives themselves are captured into the graph on their side stream)
    import torch.distributed as dist

    return "thread_local" if (dist.is_available() and dist.is_initialized()) else "global"


def _finish_hook(engine: "Engine"):
    hook = engine.grad_hook
    if hook is not None and hasattr(hook, "finish"):
        hook.finish()


class GraphedTrainStep:
    """One CUDA graph for the whole train step (forward + losses + backward + bucketed gradient all-reduce): ~370 kernel
    launches AND the NCCL collectives of `engine.grad_hook` (issued on its side stream at bucket boundaries, i.e. overlapped
    with the rest of backward) are replayed with a single cudaGraphLaunch, removing the per-launch host cost of the
    Python/ctypes orchestration at every world size.  Inputs live in static buffers; call with new tensors to copy
    them in (device-side copies)."""

    def __init__(self, engine: "Engine", x: torch.Tensor, aux: dict, train_bn: bool = True, warmup: int = 2, after_backward=None):
        self.engine = engine
        self.x = x.clone()
        self.aux = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in aux.items()}
        if self.aux.get("sym_infos") is not None:  # symmetric PM: only the (first row, count) pairs are a per-step input
            self.aux["sym_idx"] = engine.sym_table.indices(self.aux["sym_infos"])
        self.aux.pop("sym_infos", None)
        self.gw = torch.ones(8, device=x.device)

        def one_step():
            res = engine.forward(self.x, self.aux, train_bn=train_bn, do_loss=True)
            engine.backward(self.gw)
            _finish_hook(engine)
            if after_backward is not None:
                after_backward()
            return res

        snap = _bn_snapshot(engine)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # also sizes the split-K workspace and uploads the pack job table
                one_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        _bn_restore(snap)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode=_capture_mode()):
            res = one_step()
        self.losses, self.vis, self.rot, self.trans = res["losses"], res["vis"], res["rot"], res["trans"]
        self.logits = res["logits"]

    def __call__(self, x: Optional[torch.Tensor] = None, aux: Optional[dict] = None, grad_losses: Optional[torch.Tensor] = None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if aux is not None:
            if aux.get("sym_infos") is not None:
                self.aux["sym_idx"].copy_(self.engine.sym_table.indices(aux["sym_infos"]), non_blocking=True)
            for k, v in aux.items():
                if isinstance(v, torch.Tensor):
                    self.aux[k].copy_(v, non_blocking=True)
        if grad_losses is not None:
            self.gw.copy_(grad_losses)
        self.graph.replay()
        return self.losses


class GraphedFwdBwd:
    """Forward and backward captured as TWO CUDA graphs sharing one memory pool, for the public module API
    (`GDRN.forward(..., do_loss=True)` then `.backward()`): the autograd Function replays them around static buffers.
    The data-parallel all-reduces of `engine.grad_hook` are captured INSIDE the backward graph (side stream, overlapped)."""

    def __init__(self, engine: "Engine", x: torch.Tensor, aux: dict, train_bn: bool):
        self.engine, self.train_bn = engine, train_bn
        self.x = x.clone()
        self.aux = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in aux.items()}
        self.gw = torch.ones(8, device=x.device)
        self.key = self.signature(x, aux, train_bn)
        snap = _bn_snapshot(engine)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                engine.forward(self.x, self.aux, train_bn=train_bn, do_loss=True)
                engine.backward(self.gw)
                _finish_hook(engine)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        _bn_restore(snap)
        self.g_fwd, self.g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        pool = torch.cuda.graph_pool_handle()
        mode = _capture_mode()
        with torch.no_grad():
            with torch.cuda.graph(self.g_fwd, pool=pool, capture_error_mode=mode):
                res = engine.forward(self.x, self.aux, train_bn=train_bn, do_loss=True)
            with torch.cuda.graph(self.g_bwd, pool=pool, capture_error_mode=mode):
                engine.backward(self.gw)
                _finish_hook(engine)
        self.losses, self.vis = res["losses"], res["vis"]

    @staticmethod
    def signature(x, aux, train_bn):
        return (tuple(x.shape), train_bn, tuple(sorted((k, tuple(v.shape)) for k, v in aux.items() if isinstance(v, torch.Tensor))))

    def forward(self, x, aux):
        self.x.copy_(x, non_blocking=True)
        for k, v in aux.items():
            if isinstance(v, torch.Tensor):
                self.aux[k].copy_(v, non_blocking=True)
        self.g_fwd.replay()
        return self.losses.clone(), self.vis.clone()

    def backward(self, g_losses):
        self.gw.copy_(g_losses)
        self.g_bwd.replay()  # includes the bucketed all-reduces and the join of their side stream
        return self.engine.grads


class _GDRNFunction(torch.autograd.Function):
    """Autograd boundary: forward + losses and the hand-written backward are both ours; autograd only routes the
    8 loss gradients in and the 148 parameter gradients out (so optimizers / DDP / GradScaler keep working)."""

    @staticmethod
    def forward(ctx, engine: Engine, x, aux, train_bn, *params):
        ctx.engine = engine
        ctx.graphed = None
        if engine.use_cuda_graphs:
            g = engine._graphed
            if g is None or g.key != GraphedFwdBwd.signature(x, aux, train_bn):
                g = engine._graphed = GraphedFwdBwd(engine, x, aux, train_bn)
            losses, vis = g.forward(x, aux)
            ctx.graphed = g
        else:
            res = engine.forward(x, aux, train_bn=train_bn, do_loss=True)
            losses, vis = res["losses"], res["vis"]
        ctx.mark_non_differentiable(vis)
        return losses, vis

    @staticmethod
    def backward(ctx, g_losses, _g_vis):
        engine = ctx.engine
        engine.materialize_aliased_grads()
        if ctx.graphed is not None:
            ctx.graphed.backward(g_losses)
        else:
            engine.backward(g_losses)
        hook = engine.grad_hook
        if hook is not None and hasattr(hook, "finish"):
            hook.finish()  # the gradients handed to autograd are the all-reduced ones (stream-ordered wait, no host sync)
        # FRESH views of the flat gradient buffer: autograd's AccumulateGrad adopts a uniquely-referenced contiguous gradient
        # without copying, so after `zero_grad(set_to_none=True)` p.grad aliases Engine.flat_grad and the 148 per-parameter
        # clone kernels (0.4 ms per step) disappear.  materialize_aliased_grads() protects gradient accumulation.
        flat, outs, off = engine.flat_grad, [], 0
        for _, p in engine.named_params:
            n = p.numel()
            outs.append(flat[off:off + n].view_as(p))
            off += n
        return (None, None, None, None) + tuple(outs)
