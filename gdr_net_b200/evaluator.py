"""On-device pose-error metrics: ADD, ADD-S (ADI), rotation / translation error for a whole batch of
predictions in one kernel launch -- the reference evaluates them per instance on the host with numpy and a scipy KD-tree
(`lib/pysixd/pose_error.py:297-337, 400-436`, called from `core/gdrn_modeling/gdrn_evaluator.py:316-436`).

Only the metric arithmetic lives here; dataset bookkeeping, RANSAC-PnP and the BOP toolkit stay out of scope."""
from __future__ import annotations

import torch

from .capi import C


def pose_errors(R_est: torch.Tensor, t_est: torch.Tensor, R_gt: torch.Tensor, t_gt: torch.Tensor, points: torch.Tensor,
                want_adi: bool = True) -> dict:
    """R_* [B,3,3], t_* [B,3], points [B,n,3] (or [n,3] shared by the batch) on CUDA -> dict of [B] tensors `add`, `adi`, `re` (deg),
    `te` (units of t).  No host synchronisation."""
    if not R_est.is_cuda:
        raise RuntimeError("gdr_net_b200.evaluator runs on CUDA only; there is no CPU fallback")
    B = R_est.shape[0]
    if points.dim() == 2:
        points = points.unsqueeze(0).expand(B, -1, -1)
    f = lambda t: t.detach().to(R_est.device).float().contiguous()  # noqa: E731
    R_est, t_est, R_gt, t_gt, points = f(R_est), f(t_est).reshape(B, 3), f(R_gt), f(t_gt).reshape(B, 3), f(points)
    if R_gt.shape != (B, 3, 3) or points.shape[0] != B or points.shape[2] != 3:
        raise ValueError(f"pose_errors: inconsistent shapes R_gt {tuple(R_gt.shape)} points {tuple(points.shape)} for a batch of {B}")
    out = torch.empty(B, 4, device=R_est.device)
    C.gdrn_pose_errors(R_est.data_ptr(), t_est.data_ptr(), R_gt.data_ptr(), t_gt.data_ptr(), points.data_ptr(), B, points.shape[1],
                       int(want_adi), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return dict(add=out[:, 0], adi=out[:, 1], re=out[:, 2], te=out[:, 3])
