"""TEST INFRASTRUCTURE ONLY -- golden description of the host->device collate, produced by the UNMODIFIED
`core.gdrn_modeling.engine_utils.batch_data` (engine_utils.py:6-90, train and test phase) imported from /root/reference, on per-sample
dicts built from the seeded synthetic batch.  Stored per output key: dtype, shape and a SHA-1 of the tensor bytes (the inputs are
regenerated from the seed by the test).  Output: tests/golden/batch_data_b3.json.  Usage: python -m oracle.make_golden_batch_data"""
import hashlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gdr_net_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402


def per_sample_dicts(B=3, seed=4):
    """Per-sample dicts as the reference's dataset emits them (data_loader.py:617-632)."""
    b = synth.make_batch(B, seed=seed, with_sym=True)
    data = []
    for i in range(B):
        data.append(dict(roi_img=b["roi_img"][i], roi_cls=int(b["roi_cls"][i]), roi_coord_2d=b["roi_coord_2d"][i], cam=b["roi_cam"][i],
                         bbox_center=b["roi_center"][i].double(), roi_wh=b["roi_wh"][i], resize_ratio=float(b["resize_ratio"][i]),
                         roi_extent=b["roi_extent"][i], trans_ratio=b["roi_trans_ratio"][i], roi_xyz=b["roi_xyz"][i],
                         roi_mask_trunc=b["roi_mask_trunc"][i], roi_mask_visib=b["roi_mask_visib"][i], roi_mask_obj=b["roi_mask_obj"][i],
                         roi_region=b["roi_region"][i].int(), ego_rot=b["ego_rot"][i], trans=b["trans"][i], roi_points=b["roi_points"][i],
                         sym_info=b["sym_info"][i]))
    return data


def describe(batch):
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            out[k] = dict(dtype=str(v.dtype), shape=list(v.shape), sha1=hashlib.sha1(v.contiguous().numpy().tobytes()).hexdigest())
        else:
            out[k] = dict(type=type(v).__name__, len=len(v))
    return out


def main():
    ref_shim.install()
    from core.gdrn_modeling.engine_utils import batch_data

    data = per_sample_dicts()
    gold = dict(train=describe(batch_data(None, data, device="cpu")))
    path = os.path.join(ROOT, "tests", "golden", "batch_data_b3.json")
    with open(path, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", path, sorted(gold["train"]))


if __name__ == "__main__":
    main()
