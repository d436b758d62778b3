"""Test-time RANSAC-PnP on the device (the `TEST.USE_PNP` branch of the evaluator).

Reference: `GDRN_Evaluator.process_pnp_ransac` (`core/gdrn_modeling/gdrn_evaluator.py:316-436`) loops over the instances on the
host: `get_out_coor` / `get_out_mask` (`engine_utils.py:92-126`), D2H copies, `get_img_model_points_with_coords2d` (:89-126) and
`misc.pnp_v2(..., cv2.SOLVEPNP_EPNP, ransac=True, ransac_reprojErr=3, ransac_iter=100)` (`lib/pysixd/misc.py:145-194`), then
replaces the translation by the one decoded from the network's (cx, cy, z) head (:398-421).

Here the whole batch is solved by three kernel launches (`csrc/pnp_ransac.cu`), following OpenCV's algorithm closely enough that
the inlier sets and poses agree with cv2 up to round-off (tests/test_pnp_ransac_gpu.py).  No host synchronisation."""
from __future__ import annotations

import torch

from .capi import C

_MASK_MODES = {"none": 0, "L1": 1, "BCE": 2}


def pnp_ransac(mask: torch.Tensor, xyz: torch.Tensor, coord_2d: torch.Tensor, extents: torch.Tensor, im_wh, cams: torch.Tensor,
               mask_loss_type: str = "L1", mask_thr: float = 0.5, reproj_err: float = 3.0, iters: int = 100,
               confidence: float = 0.99, return_inliers: bool = False) -> dict:
    """mask [B,1,h,w] raw head output (normalised here like `get_out_mask` for `mask_loss_type` "L1" / "BCE"; "none": used as is),
    xyz [B,3,h,w] in [0,1] (`get_out_coor` of the regression head = cat(coor_x, coor_y, coor_z)), coord_2d [B,2,h,w] in [0,1],
    extents [B,3], im_wh [B,2] or (W, H), cams [B,3,3] or [3,3].
    Returns dict(pose [B,3,4], num_points [B], num_inliers [B], iterations [B], ok [B] (bool)[, inliers [B,h*w] uint8 flags over
    the row-major list of selected points]).  Instances with fewer than 5 selected points get ok = False and an identity pose
    (the reference writes -100 there, gdrn_evaluator.py:388-390; `process_pnp_ransac` below does the same)."""
    if not mask.is_cuda:
        raise RuntimeError("gdr_net_b200.pnp_ransac runs on CUDA only; there is no CPU fallback")
    if mask_loss_type not in _MASK_MODES:
        raise NotImplementedError(f"unknown mask loss type: {mask_loss_type}")  # "CE" masks are not produced by the shipped configs
    dev = mask.device
    B, _, h, w = xyz.shape
    f = lambda t: t.detach().to(dev).float().contiguous()  # noqa: E731
    mask, xyz, coord_2d, extents = f(mask).reshape(B, h * w), f(xyz), f(coord_2d), f(extents).reshape(B, 3)
    if coord_2d.shape != (B, 2, h, w):
        raise ValueError(f"pnp_ransac: coord_2d {tuple(coord_2d.shape)} does not match xyz {tuple(xyz.shape)}")
    if not torch.is_tensor(im_wh):
        im_wh = torch.tensor(im_wh, dtype=torch.float32)
    im_wh = f(im_wh)
    im_wh = (im_wh.reshape(1, 2).expand(B, 2) if im_wh.numel() == 2 else im_wh.reshape(B, 2)).contiguous()
    cams = f(cams)
    cams = (cams.reshape(1, 3, 3).expand(B, 3, 3) if cams.dim() == 2 else cams.reshape(B, 3, 3)).contiguous()
    nbytes = int(C.load().gdrn_pnp_ransac_workspace_bytes(B, h * w, iters))
    ws = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
    pose = torch.empty(B, 3, 4, device=dev)
    info = torch.empty(B, 4, dtype=torch.int32, device=dev)
    inl = torch.empty(B, h * w, dtype=torch.uint8, device=dev) if return_inliers else None
    C.gdrn_pnp_ransac(mask.data_ptr(), xyz.data_ptr(), coord_2d.data_ptr(), extents.data_ptr(), im_wh.data_ptr(), cams.data_ptr(),
                      B, h, w, _MASK_MODES[mask_loss_type], float(mask_thr), float(reproj_err), int(iters), float(confidence),
                      ws.data_ptr(), ws.numel() * 8, pose.data_ptr(), info.data_ptr(), inl.data_ptr() if inl is not None else None,
                      torch.cuda.current_stream().cuda_stream)
    out = dict(pose=pose, num_points=info[:, 0], num_inliers=info[:, 1], iterations=info[:, 2], ok=info[:, 3] != 0)
    if inl is not None:
        out["inliers"] = inl
    return out


def process_pnp_ransac(out_dict: dict, roi_coord_2d: torch.Tensor, roi_extents: torch.Tensor, im_wh, cams: torch.Tensor,
                       mask_loss_type: str = "L1", mask_thr: float = 0.5, use_net_trans: bool = True) -> torch.Tensor:
    """Batched equivalent of the pose part of `process_pnp_ransac` (gdrn_evaluator.py:316-421): rotation from RANSAC-PnP on the
    predicted coordinate maps, translation from the network (`out_dict["trans"]`, already decoded on the device by the model's
    test branch) when `use_net_trans`.  out_dict: the model's test-mode output with `TEST.USE_PNP` keys (mask, coor_x/y/z, trans).
    Returns [B,3,4]; instances without enough points get the reference's -100 marker."""
    xyz = torch.cat([out_dict["coor_x"], out_dict["coor_y"], out_dict["coor_z"]], dim=1)
    res = pnp_ransac(out_dict["mask"], xyz, roi_coord_2d, roi_extents, im_wh, cams, mask_loss_type=mask_loss_type, mask_thr=mask_thr)
    pose = res["pose"].clone()
    pose = torch.where(res["ok"].view(-1, 1, 1), pose, torch.full_like(pose, -100.0))
    if use_net_trans:
        pose[:, :, 3] = out_dict["trans"].to(pose.dtype)
    return pose
