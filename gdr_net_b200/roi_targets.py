"""Per-instance crop / target generation on the GPU: what the reference's dataloader workers do with
cv2 / scipy for every training instance (`core/gdrn_modeling/data_loader.py:487-560`, `core/utils/data_utils.py:80-137,213-219`)
as two kernel launches per batch.  Inputs are the full-resolution arrays the reference has right before its
`crop_resize_by_warp_affine` calls; outputs carry the reference's batch keys, already on the device."""
from __future__ import annotations

import torch

from .capi import C


def _f(t, dev):
    return torch.as_tensor(t).to(dev).float().contiguous()


def make_roi_batch(image_u8: torch.Tensor, xyz: torch.Tensor, mask_visib: torch.Tensor, mask_trunc, bbox_center: torch.Tensor,
                   scale: torch.Tensor, extents: torch.Tensor, fps_points: torch.Tensor, input_res: int = 256, out_res: int = 64,
                   pixel_std: float = 255.0) -> dict:
    """image_u8 [B,H,W,3] uint8 (BGR, cfg.INPUT.FORMAT), xyz [B,H,W,3] float (0 outside the object), mask_visib [B,H,W] (the visible
    segmentation), mask_trunc [B,H,W] or None, bbox_center [B,2], scale [B] (the augmented square crop side), extents [B,3],
    fps_points [B,F,3] (F = NUM_REGIONS farthest-point-sampled model points) -> dict with `roi_img` [B,3,256,256] (/255, PIXEL_STD),
    `roi_coord_2d`, `roi_xyz`, `roi_mask_trunc`, `roi_mask_visib`, `roi_mask_obj`, `roi_region` (int64), `resize_ratio`."""
    if not image_u8.is_cuda:
        raise RuntimeError("gdr_net_b200.roi_targets runs on CUDA only; there is no CPU fallback")
    dev = image_u8.device
    if image_u8.dtype != torch.uint8 or image_u8.dim() != 4 or image_u8.shape[-1] != 3:
        raise ValueError(f"image_u8 must be uint8 [B,H,W,3], got {image_u8.dtype} {tuple(image_u8.shape)}")
    B, H, W, _ = image_u8.shape
    image_u8 = image_u8.contiguous()
    xyz, mask_visib = _f(xyz, dev), _f(mask_visib, dev)
    mask_trunc = None if mask_trunc is None else _f(mask_trunc, dev)
    # bbox centre / scale stay float64 like the reference's aug_bbox output: get_affine_transform rounds them to float32 at specific
    # points of its computation and the kernels reproduce exactly those roundings
    centers = torch.as_tensor(bbox_center).to(dev).double().reshape(B, 2).contiguous()
    scales = torch.as_tensor(scale).to(dev).double().reshape(B).contiguous()
    extents, fps = _f(extents, dev).reshape(B, 3), _f(fps_points, dev)
    if xyz.shape != (B, H, W, 3) or mask_visib.shape != (B, H, W) or fps.dim() != 3 or fps.shape[0] != B or fps.shape[2] != 3:
        raise ValueError(f"make_roi_batch: inconsistent shapes: image {tuple(image_u8.shape)}, xyz {tuple(xyz.shape)}, mask_visib "
                         f"{tuple(mask_visib.shape)}, fps_points {tuple(fps.shape)}")
    if mask_trunc is not None and mask_trunc.shape != (B, H, W):
        raise ValueError(f"make_roi_batch: mask_trunc {tuple(mask_trunc.shape)} does not match the images [{B},{H},{W}]")
    if fps.shape[1] < 1 or fps.shape[1] > 1024:
        raise ValueError(f"make_roi_batch: {fps.shape[1]} region anchors (1..1024 supported)")
    s = torch.cuda.current_stream().cuda_stream
    out = dict(roi_img=torch.empty(B, 3, input_res, input_res, device=dev),
               roi_xyz=torch.empty(B, 3, out_res, out_res, device=dev),
               roi_mask_trunc=torch.empty(B, out_res, out_res, device=dev), roi_mask_visib=torch.empty(B, out_res, out_res, device=dev),
               roi_mask_obj=torch.empty(B, out_res, out_res, device=dev),
               roi_region=torch.empty(B, out_res, out_res, dtype=torch.long, device=dev),
               roi_coord_2d=torch.empty(B, 2, out_res, out_res, device=dev))
    C.gdrn_roi_crop_image(image_u8.data_ptr(), centers.data_ptr(), scales.data_ptr(), out["roi_img"].data_ptr(), B, H, W, input_res,
                          float(pixel_std), s)
    C.gdrn_roi_targets(xyz.data_ptr(), mask_visib.data_ptr(), None if mask_trunc is None else mask_trunc.data_ptr(), centers.data_ptr(),
                       scales.data_ptr(), extents.data_ptr(), fps.data_ptr(), fps.shape[1], out["roi_xyz"].data_ptr(),
                       out["roi_mask_trunc"].data_ptr(), out["roi_mask_visib"].data_ptr(), out["roi_mask_obj"].data_ptr(),
                       out["roi_region"].data_ptr(), out["roi_coord_2d"].data_ptr(), B, H, W, out_res, s)
    out["resize_ratio"] = (out_res / scales).float()  # data_loader.py:621
    return out
