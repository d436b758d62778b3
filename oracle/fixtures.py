"""TEST INFRASTRUCTURE ONLY -- shared seeded fixtures for oracle / parity tests.

`calibrated_state_dict` = `synth.seeded_state_dict` (Kaiming-scale weights) whose
BatchNorm running statistics are then set from one oracle train-mode pass over a seeded
calibration batch and jittered, so that eval-mode activations stay O(1) through all 50 layers
(random running stats make an un-normalised ResNet whose activations grow to ~1e4, which
saturates the region softmax and makes relative-error tests meaningless).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from gdr_net_b200 import synth
from oracle import gdrn_oracle as O

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
_CACHE: dict = {}


def template_from_manifest() -> dict:
    m = np.load(os.path.join(_GOLDEN, "state_dict_manifest.npz"))
    out = {}
    for name, shape in zip(m["names"], m["shapes"]):
        shape = eval(str(shape))
        dt = torch.long if str(name).endswith("num_batches_tracked") else torch.float32
        out[str(name)] = torch.zeros(shape, dtype=dt)
    return out


def calibrated_state_dict(seed: int = 0, template: dict | None = None) -> dict:
    key = (seed, id(template) if template is not None else None)
    if key in _CACHE:
        return {k: v.clone() for k, v in _CACHE[key].items()}
    template = template if template is not None else template_from_manifest()
    sd = synth.seeded_state_dict(template, seed)
    work = O.leaf_state_dict(sd, requires_grad=False)
    batch = synth.make_batch(4, seed=seed + 4321)
    O.BN_MOMENTUM[0] = 1.0
    try:
        with torch.no_grad():
            feat = O.backbone_forward(batch["roi_img"], work, train=True, update_stats=True)
            O.head_forward(feat, work, train=True, update_stats=True)
    finally:
        O.BN_MOMENTUM[0] = 0.1
    for k in sd:
        g = synth._gen(seed, "calib/" + k)
        if k.endswith("running_mean"):
            var = work[k.replace("running_mean", "running_var")]
            sd[k] = work[k] + 0.1 * var.sqrt() * torch.randn(var.shape, generator=g)
        elif k.endswith("running_var"):
            sd[k] = work[k] * (0.8 + 0.45 * torch.rand(work[k].shape, generator=g))
    _CACHE[key] = {k: v.clone() for k, v in sd.items()}
    return sd
