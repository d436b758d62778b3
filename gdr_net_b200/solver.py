"""Optimizer construction for `build_model_optimizer` (reference core/utils/solver_utils.py:17-57).

The reference resolves `cfg.SOLVER.OPTIMIZER_CFG.type` through the mmcv OPTIMIZERS registry; the shipped
configs use "Ranger" (lib/torch_utils/solver/ranger.py:100-200: RAdam + Lookahead(k=6, alpha=0.5) + gradient
centralisation).  This is an independent implementation of that published algorithm:
  * CUDA parameters: ONE hand-written kernel launch per param group (csrc/optim.cu `gdrn_ranger_step`: centralisation +
    RAdam + lookahead over a device-resident job table), 28 B of HBM traffic per parameter;
  * CPU parameters (unit tests of the algorithm): a multi-tensor `torch._foreach` restatement.
State keys (`step`, `exp_avg`, `exp_avg_sq`, `slow_buffer`) are the reference's, so optimizer checkpoints interchange.
"""
from __future__ import annotations

import math

import torch


class Ranger(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(0.95, 0.999), eps=1e-5,
                 weight_decay=0, use_gc=True, gc_conv_only=False, fused=True):
        if not 0.0 <= alpha <= 1.0:
            raise ValueError(f"Invalid slow update rate: {alpha}")
        if not 1 <= k:
            raise ValueError(f"Invalid lookahead steps: {k}")
        if not lr > 0:
            raise ValueError(f"Invalid Learning Rate: {lr}")
        if not eps > 0:
            raise ValueError(f"Invalid eps: {eps}")
        defaults = dict(lr=lr, alpha=alpha, k=k, step_counter=0, betas=betas, N_sma_threshhold=N_sma_threshhold, eps=eps,
                        weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.N_sma_threshhold = N_sma_threshhold
        self.alpha = alpha
        self.k = k
        self.use_gc = use_gc
        self.gc_gradient_threshold = 3 if gc_conv_only else 1
        self.fused = fused      # CUDA fp32 parameters: one hand-written kernel per param group (csrc/optim.cu)
        self.grad_scale = 1.0   # multiplies every gradient inside the step (e.g. 1/loss-scale, 1/accumulation steps)

    # ------------------------------------------------------------------------------------------ fused CUDA path
    def _fused_tables(self, gi, ps):
        """Device job / block tables of one param group (csrc/optim.cu), rebuilt only when a tensor address changes."""
        import numpy as np

        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        cache = self.__dict__.setdefault("_fused_cache", {})
        ent = cache.get(gi)
        if ent is not None and ent["key"] == key:
            return ent
        jdt = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("slow", "<u8"), ("numel", "<i8"),
                        ("row_len", "<i4"), ("pad", "<i4", (3,))])
        bdt = np.dtype([("job", "<i4"), ("count", "<i4"), ("begin", "<i8")])
        assert jdt.itemsize == 64 and bdt.itemsize == 16
        jobs = np.zeros(len(ps), dtype=jdt)
        blocks = []
        for j, p in enumerate(ps):
            st = self.state[p]
            n = p.numel()
            gc = self.use_gc and p.dim() > self.gc_gradient_threshold
            row_len = n // p.shape[0] if gc else 0
            jobs[j] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                       st["slow_buffer"].data_ptr(), n, row_len, (0, 0, 0))
            if row_len:
                rows = p.shape[0]
                per = max(8, 8192 // row_len) if row_len <= 2048 else 1  # <= 2048: one warp per row, 8 rows in flight
                for r0 in range(0, rows, per):
                    blocks.append((j, min(per, rows - r0), r0))
            else:
                for e0 in range(0, n, 8192):
                    blocks.append((j, min(8192, n - e0), e0))
        dev = ps[0].device
        ent = dict(key=key, jobs=torch.from_numpy(jobs.view(np.uint8).copy()).to(dev),
                   blocks=torch.from_numpy(np.array(blocks, dtype=bdt).view(np.uint8).copy()).to(dev), nblocks=len(blocks))
        cache[gi] = ent
        return ent

    def _fused_group_step(self, gi, group, ps, step):
        from .capi import C

        beta1, beta2 = group["betas"]
        n_sma, step_size = self._radam(step, beta1, beta2)
        ent = self._fused_tables(gi, ps)
        C.gdrn_ranger_step(ent["jobs"].data_ptr(), ent["blocks"].data_ptr(), ent["nblocks"], float(step_size * group["lr"]),
                           float(group["weight_decay"] * group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                           float(self.alpha), float(self.grad_scale), int(n_sma > self.N_sma_threshhold),
                           int(step % group["k"] == 0), torch.cuda.current_stream().cuda_stream)

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
            st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
            st["slow_buffer"] = p.detach().clone()
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        for gi, group in enumerate(self.param_groups):
            beta1, beta2 = group["betas"]
            lr, eps, wd = group["lr"], group["eps"], group["weight_decay"]
            live = [p for p in group["params"] if p.grad is not None]
            if self.fused and live and all(p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32
                                           and p.is_contiguous() and p.grad.is_contiguous() for p in live):
                steps = {self._init_state(p)["step"] for p in live}
                if len(steps) == 1:  # the usual case: one launch for the whole group
                    step = steps.pop() + 1
                    for p in live:
                        self.state[p]["step"] = step
                    self._fused_group_step(gi, group, live, step)
                    continue
            ps, gs, m1, m2, slow = [], [], [], [], []
            step = None
            for p in live:
                g = p.grad.float()
                if self.grad_scale != 1.0:
                    g = g * self.grad_scale
                st = self._init_state(p)
                if self.use_gc and g.dim() > self.gc_gradient_threshold:
                    g = g - g.mean(dim=tuple(range(1, g.dim())), keepdim=True)
                st["step"] += 1
                step = st["step"] if step is None else step
                if st["step"] != step:  # parameters joined the group at different times: fall back per tensor
                    self._step_one(p, g, st, group)
                    continue
                ps.append(p)
                gs.append(g)
                m1.append(st["exp_avg"])
                m2.append(st["exp_avg_sq"])
                slow.append(st["slow_buffer"])
            if not ps:
                continue
            torch._foreach_mul_(m2, beta2)
            torch._foreach_addcmul_(m2, gs, gs, value=1 - beta2)
            torch._foreach_mul_(m1, beta1)
            torch._foreach_add_(m1, gs, alpha=1 - beta1)
            n_sma, step_size = self._radam(step, beta1, beta2)
            if wd != 0:
                torch._foreach_mul_(ps, 1 - wd * lr)
            if n_sma > self.N_sma_threshhold:
                denom = torch._foreach_sqrt(m2)
                torch._foreach_add_(denom, eps)
                torch._foreach_addcdiv_(ps, m1, denom, value=-step_size * lr)
            else:
                torch._foreach_add_(ps, m1, alpha=-step_size * lr)
            if step % group["k"] == 0:
                diff = torch._foreach_sub(ps, slow)
                torch._foreach_add_(slow, diff, alpha=self.alpha)
                for p, s in zip(ps, slow):
                    p.copy_(s)
        return loss

    def _radam(self, step, beta1, beta2):
        beta2_t = beta2 ** step
        n_sma_max = 2 / (1 - beta2) - 1
        n_sma = n_sma_max - 2 * step * beta2_t / (1 - beta2_t)
        if n_sma > self.N_sma_threshhold:
            step_size = math.sqrt(
                (1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max / (n_sma_max - 2)
            ) / (1 - beta1 ** step)
        else:
            step_size = 1.0 / (1 - beta1 ** step)
        return n_sma, step_size

    def _step_one(self, p, g, st, group):
        beta1, beta2 = group["betas"]
        st["exp_avg_sq"].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        st["exp_avg"].mul_(beta1).add_(g, alpha=1 - beta1)
        n_sma, step_size = self._radam(st["step"], beta1, beta2)
        if group["weight_decay"] != 0:
            p.mul_(1 - group["weight_decay"] * group["lr"])
        if n_sma > self.N_sma_threshhold:
            p.addcdiv_(st["exp_avg"], st["exp_avg_sq"].sqrt().add_(group["eps"]), value=-step_size * group["lr"])
        else:
            p.add_(st["exp_avg"], alpha=-step_size * group["lr"])
        if st["step"] % group["k"] == 0:
            st["slow_buffer"].add_(p - st["slow_buffer"], alpha=self.alpha)
            p.copy_(st["slow_buffer"])


OPTIMIZERS = {"Ranger": Ranger, "SGD": torch.optim.SGD, "Adam": torch.optim.Adam, "AdamW": torch.optim.AdamW,
              "RMSprop": torch.optim.RMSprop}


def build_optimizer_with_params(cfg, params):
    opt_cfg = cfg.SOLVER.OPTIMIZER_CFG
    if opt_cfg == "":
        raise RuntimeError("please provide cfg.SOLVER.OPTIMIZER_CFG to build optimizer")
    if isinstance(opt_cfg, str):
        opt_cfg = eval(opt_cfg)
    args = dict(opt_cfg)
    typ = args.pop("type")
    if typ not in OPTIMIZERS:
        raise ValueError(f"Unknown optimizer name: {typ}")
    return OPTIMIZERS[typ](params, **args)
