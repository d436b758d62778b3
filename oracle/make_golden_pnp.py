"""TEST INFRASTRUCTURE ONLY -- golden vectors for the evaluator's RANSAC-PnP path, produced by the UNMODIFIED
reference functions imported from /root/reference (run in the build container; the GPU box has no /root/reference):

  * `core.gdrn_modeling.engine_utils.get_out_mask`                       (engine_utils.py:108-126, L1 mask head)
  * `GDRN_Evaluator.get_img_model_points_with_coords2d`                  (gdrn_evaluator.py:89-126)
  * `lib.pysixd.misc.pnp_v2(..., cv2.SOLVEPNP_EPNP, ransac=True, 3, 100)` (misc.py:145-194), exactly as gdrn_evaluator.py:372-387

on the seeded synthetic head outputs of `gdr_net_b200.synth.make_pnp_maps`.  Output: tests/golden/pnp_ransac_b4.npz (inputs,
the selected image / model points per ROI, the [3,4] poses).  Usage: python -m oracle.make_golden_pnp"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gdr_net_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402


def _stub_plotting():
    """The evaluator module imports visualisation helpers (matplotlib) at module level; none of them is on this path."""
    class _Any:
        def __getattr__(self, n):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

        def __setitem__(self, k, v):
            pass

        def __getitem__(self, k):
            return _Any()

    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d"]:
        m = types.ModuleType(name)
        m.__path__ = []
        m.rcParams = {}
        m.__getattr__ = lambda n: _Any()
        sys.modules[name] = m


def main():
    import cv2

    ref_shim.install()
    _stub_plotting()
    import core.gdrn_modeling.engine_utils as engine_utils
    import core.gdrn_modeling.gdrn_evaluator as evaluator
    import lib.pysixd.misc as misc

    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(CDPN=types.SimpleNamespace(ROT_HEAD=types.SimpleNamespace(MASK_LOSS_TYPE="L1"))))
    d = synth.make_pnp_maps(4, seed=7)
    out_mask = engine_utils.get_out_mask(cfg, d["mask"]).numpy()
    out_xyz = engine_utils.get_out_coor(cfg, d["xyz"][:, 0:1], d["xyz"][:, 1:2], d["xyz"][:, 2:3]).numpy()
    save = {k: v.numpy() for k, v in d.items()}
    for b in range(4):
        im_W, im_H = int(d["im_wh"][b, 0]), int(d["im_wh"][b, 1])
        xyz_i = out_xyz[b].transpose(1, 2, 0).copy()
        mask_i = np.squeeze(out_mask[b])
        coord_2d_i = d["coord_2d"][b].numpy().transpose(1, 2, 0).copy()
        img_points, model_points = evaluator.GDRN_Evaluator.get_img_model_points_with_coords2d(
            None, mask_i, xyz_i, coord_2d_i, im_H=im_H, im_W=im_W, extent=d["extents"][b].numpy(), mask_thr=0.5)
        K = d["cams"][b].numpy().copy()
        pose = misc.pnp_v2(model_points, img_points, K, method=cv2.SOLVEPNP_EPNP, ransac=True, ransac_reprojErr=3, ransac_iter=100)
        save[f"img_points_{b}"] = img_points
        save[f"model_points_{b}"] = model_points
        save[f"pose_{b}"] = np.asarray(pose, np.float64)
        print(b, len(img_points), pose[:, 3])
    save["cv2_version"] = np.array(cv2.__version__)
    out = os.path.join(ROOT, "tests", "golden", "pnp_ransac_b4.npz")
    np.savez_compressed(out, **save)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
