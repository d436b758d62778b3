"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-instance crop / target generation
(core/gdrn_modeling/data_loader.py:487-560) with the SAME libraries the reference uses (cv2.warpAffine, scipy cdist), for the
parity test of gdr_net_b200.roi_targets.  Citations are relative to /root/reference/.

Pinned: oracle/make_golden_roi.py imports the UNMODIFIED helpers (`crop_resize_by_warp_affine`, `get_2d_coord_np`, `xyz_to_region`)
from /root/reference and stores their outputs in tests/golden/roi_targets_b3.npz; the restatements below reproduce them bit for bit
(tests/test_roi_targets_cpu.py::test_roi_oracle_matches_reference_golden)."""
from __future__ import annotations

import cv2
import numpy as np
from scipy.spatial.distance import cdist


def _get_dir(src_point, rot_rad):
    """core/utils/data_utils.py:151-158"""
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    src_result = [0, 0]
    src_result[0] = src_point[0] * cs - src_point[1] * sn
    src_result[1] = src_point[0] * sn + src_point[1] * cs
    return src_result


def _get_3rd_point(a, b):
    """core/utils/data_utils.py:146-148"""
    direct = a - b
    return b + np.array([-direct[1], direct[0]], dtype=np.float32)


def get_affine_transform(center, scale, output_size):
    """core/utils/data_utils.py:96-137 with rot = 0, shift = 0, inv = False; `scale` = (s, s) as crop_resize_by_warp_affine passes it.
    Operand types are kept as in the reference (float64 centre / scale, float32 point arrays): their roundings matter."""
    scale_tmp = (scale, scale)
    shift = np.array([0, 0], dtype=np.float32)
    src_w, dst_w, dst_h = scale_tmp[0], output_size, output_size
    src_dir = _get_dir([0, src_w * -0.5], np.pi * 0 / 180)
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir
    src[2:, :] = _get_3rd_point(src[0, :], src[1, :])
    dst[2:, :] = _get_3rd_point(dst[0, :], dst[1, :])
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def crop_resize_by_warp_affine(img, center, scale, output_size, interpolation):
    """core/utils/data_utils.py:80-93"""
    trans = get_affine_transform(center, scale, output_size)
    return cv2.warpAffine(img, trans, (int(output_size), int(output_size)), flags=interpolation)


def xyz_to_region(xyz_crop, fps_points):
    """core/utils/data_utils.py:213-219"""
    bh, bw = xyz_crop.shape[:2]
    mask_crop = ((xyz_crop[:, :, 0] != 0) | (xyz_crop[:, :, 1] != 0) | (xyz_crop[:, :, 2] != 0)).astype("uint8")
    dists = cdist(xyz_crop.reshape(bh * bw, 3), fps_points)
    region_ids = np.argmin(dists, axis=1).reshape(bh, bw) + 1
    return mask_crop * region_ids


def roi_instance(image, xyz, seg_visib, mask_trunc, bbox_center, scale, extent, fps_points, input_res=256, out_res=64):
    """data_loader.py:472 (mask_obj), :487-545: one training instance."""
    H, W = image.shape[:2]
    mask_obj = ((xyz[:, :, 0] != 0) | (xyz[:, :, 1] != 0) | (xyz[:, :, 2] != 0)).astype(np.float32)
    roi_img = crop_resize_by_warp_affine(image, bbox_center, scale, input_res, cv2.INTER_LINEAR).transpose(2, 0, 1)
    roi_img = roi_img.astype(np.float32) / 255.0  # normalize_image with PIXEL_MEAN 0 / PIXEL_STD 255 (gdrn_base.py:12-13)
    x = np.linspace(0, 1, W, dtype=np.float32)
    y = np.linspace(0, 1, H, dtype=np.float32)
    coord_2d = np.asarray(np.meshgrid(x, y)).transpose(1, 2, 0)  # get_2d_coord_np(..., fmt="HWC")
    roi_coord_2d = crop_resize_by_warp_affine(coord_2d, bbox_center, scale, out_res, cv2.INTER_LINEAR).transpose(2, 0, 1)
    mask_visib = seg_visib.astype("float32") * mask_obj
    m_trunc = mask_visib if mask_trunc is None else mask_visib * mask_trunc.astype("float32")
    near = cv2.INTER_NEAREST
    roi_mask_trunc = crop_resize_by_warp_affine(m_trunc[:, :, None], bbox_center, scale, out_res, near)
    roi_mask_visib = crop_resize_by_warp_affine(mask_visib[:, :, None], bbox_center, scale, out_res, near)
    roi_mask_obj = crop_resize_by_warp_affine(mask_obj[:, :, None], bbox_center, scale, out_res, near)
    roi_xyz = crop_resize_by_warp_affine(xyz, bbox_center, scale, out_res, near)
    roi_region = xyz_to_region(roi_xyz, fps_points)
    roi_xyz = roi_xyz.transpose(2, 0, 1).copy()
    for c in range(3):
        roi_xyz[c] = roi_xyz[c] / extent[c] + 0.5
    return dict(roi_img=roi_img, roi_coord_2d=roi_coord_2d.astype(np.float32), roi_mask_trunc=roi_mask_trunc, roi_mask_visib=roi_mask_visib,
                roi_mask_obj=roi_mask_obj, roi_xyz=roi_xyz.astype(np.float32), roi_region=roi_region.astype(np.int64),
                resize_ratio=out_res / scale)
