"""TEST INFRASTRUCTURE ONLY -- golden vectors for the `PNP_NET.WITH_2D_COORD = False` variant (Patch-PnP input = xyz + 64 region
channels, nIn = 67), produced by the UNMODIFIED reference model (GDRN.py:156-169, 635-647) built from the a6
config with that one switch changed: train-mode forward + losses + backward on a seeded batch of FOUR crops.
(Not three: `allo_to_ego_mat_torch` (core/utils/utils.py:219) calls `torch.cross` without `dim`, which in the installed torch still means
"the first dimension of size 3" -- for a batch of exactly 3 crops the reference crosses along the BATCH dimension.  That accident
of the deprecated default is not reproduced here or in the CUDA path; every other batch size takes dim = 1.)
Output: tests/golden/train_nin67_b4.npz.  Usage: python -m oracle.make_golden_nin67"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gdr_net_b200 import synth  # noqa: E402
from gdr_net_b200.config import Config, postprocess_like_main_gdrn  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import REF_CFG  # noqa: E402

GRAD_KEYS = ("pnp_net.features.0.weight", "pnp_net.fc_r.weight", "rot_head_net.features.23.bias", "backbone.bn1.weight")


def main():
    torch.set_num_threads(os.cpu_count())
    ref_gdrn = ref_shim.import_reference_gdrn()
    cfg = Config.fromfile(os.path.join(ref_shim.REFERENCE_ROOT, REF_CFG))
    cfg = postprocess_like_main_gdrn(cfg, device="cpu")
    cfg.MODEL.CDPN.BACKBONE.PRETRAINED = ""
    cfg.MODEL.CDPN.PNP_NET.WITH_2D_COORD = False
    torch.manual_seed(0)
    model, _ = ref_gdrn.build_model_optimizer(copy.deepcopy(cfg))
    assert model.pnp_net.features[0].in_channels == 67
    sd = synth.seeded_state_dict(model.state_dict(), seed=3)
    model.load_state_dict(sd)
    model.train()
    batch = synth.make_batch(4, seed=17)  # NOT 3: see the note in the module docstring
    _, loss_dict = model(batch["roi_img"].clone(), **synth.forward_kwargs(batch, train=True))
    sum(loss_dict.values()).backward()
    save = {("loss/" + k): np.array(float(v)) for k, v in loss_dict.items()}
    grads = dict(model.named_parameters())
    for k in GRAD_KEYS:
        save["grad/" + k] = grads[k].grad.detach().numpy()
    path = os.path.join(ROOT, "tests", "golden", "train_nin67_b4.npz")
    np.savez_compressed(path, **save)
    print({k: float(v) for k, v in loss_dict.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
