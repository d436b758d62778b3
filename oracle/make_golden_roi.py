"""TEST INFRASTRUCTURE ONLY -- golden vectors for the crop / target generation path, produced by the UNMODIFIED
reference helpers imported from /root/reference (run in the build container):

  * `core.utils.data_utils.crop_resize_by_warp_affine`  (data_utils.py:80-93, -> get_affine_transform :96-137 -> cv2.warpAffine)
  * `core.utils.data_utils.get_2d_coord_np`             (the [0,1] coordinate grid the loader crops, data_loader.py:346)
  * `core.utils.data_utils.xyz_to_region`               (data_utils.py:213-219)

called the way `data_loader.py:487-545` calls them, on seeded random instances.  Output: tests/golden/roi_targets_b3.npz (inputs and
the helpers' raw outputs).  oracle/roi_oracle.py restates these helpers; tests/test_roi_targets_cpu.py compares the two bit for bit.
Usage: python -m oracle.make_golden_roi"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402


def make_inputs(seed=11, B=3, H=120, W=160, F_=64):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    xyz = np.zeros((B, H, W, 3), np.float32)
    seg = np.zeros((B, H, W), np.float32)
    trunc = (rng.random((B, H, W)) > 0.2).astype(np.float32)
    centers = np.stack([rng.uniform(50, 110, B), rng.uniform(40, 80, B)], 1)
    centers[B - 1] = [7.3, 5.7]  # crop hanging over the border
    scales = rng.uniform(25, 75, B)
    ext = rng.uniform(0.05, 0.3, (B, 3)).astype(np.float32)
    fps = ((rng.random((B, F_, 3)) - 0.5) * ext[:, None, :]).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        blob = ((xx - centers[b, 0]) ** 2 + (yy - centers[b, 1]) ** 2) < (0.35 * scales[b]) ** 2
        xyz[b][blob] = ((rng.random((int(blob.sum()), 3)) - 0.5) * ext[b]).astype(np.float32)
        seg[b] = (blob & (rng.random((H, W)) > 0.1)).astype(np.float32)
    return dict(image=img, xyz=xyz, seg=seg, trunc=trunc, centers=centers, scales=scales, extents=ext, fps=fps)


def main():
    import cv2

    ref_shim.install()
    import core.utils.data_utils as du

    inp = make_inputs()
    B, H, W = inp["image"].shape[:3]
    in_res, out_res = 128, 64
    save = dict(inp)
    save["in_res"], save["out_res"] = np.array(in_res), np.array(out_res)
    coord_2d = du.get_2d_coord_np(W, H, fmt="HWC")
    for b in range(B):
        c, s = inp["centers"][b], float(inp["scales"][b])
        xyz = inp["xyz"][b]
        mask_obj = ((xyz[:, :, 0] != 0) | (xyz[:, :, 1] != 0) | (xyz[:, :, 2] != 0)).astype(np.float32)
        save[f"img_{b}"] = du.crop_resize_by_warp_affine(inp["image"][b], c, s, in_res, interpolation=cv2.INTER_LINEAR)
        save[f"coord_{b}"] = du.crop_resize_by_warp_affine(coord_2d, c, s, out_res, interpolation=cv2.INTER_LINEAR)
        save[f"obj_{b}"] = du.crop_resize_by_warp_affine(mask_obj[:, :, None], c, s, out_res, interpolation=cv2.INTER_NEAREST)
        roi_xyz = du.crop_resize_by_warp_affine(xyz, c, s, out_res, interpolation=cv2.INTER_NEAREST)
        save[f"xyz_{b}"] = roi_xyz
        save[f"region_{b}"] = du.xyz_to_region(roi_xyz, inp["fps"][b])
    save["cv2_version"] = np.array(cv2.__version__)
    out = os.path.join(ROOT, "tests", "golden", "roi_targets_b3.npz")
    np.savez_compressed(out, **save)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
