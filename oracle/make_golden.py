"""TEST INFRASTRUCTURE ONLY -- generates `tests/golden/*.npz` by running the UNMODIFIED
reference (`/root/reference`, imported through `oracle/ref_shim.py`) on seeded synthetic
inputs, and checks the oracle restatement (`oracle/gdrn_oracle.py`) against it.

Run in the build container only (the reference is absent on the GPU box):

    python -m oracle.make_golden            # writes tests/golden/, prints max deviations

The reference ships no golden vectors of its own; these
fixtures are "outputs of the reference itself run here".
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gdr_net_b200 import synth  # noqa: E402
from gdr_net_b200.config import Config, postprocess_like_main_gdrn  # noqa: E402
from oracle import gdrn_oracle as O  # noqa: E402
from oracle import fixtures, ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REF_CFG = "configs/gdrn/lm/a6_cPnP_lm13.py"
REF_CFG_YCBV = "configs/gdrn/ycbv/a6_cPnP_AugAAETrunc_BG0.5_Rsym_ycbv_real_pbr_visib20_10e.py"

# parameters whose full gradients are stored (small ones) -- the rest are stored as norms + strided samples
FULL_GRAD = ("pnp_net.fc_r.weight", "pnp_net.fc_t.weight", "pnp_net.fc_r.bias", "backbone.bn1.weight",
             "backbone.bn1.bias", "rot_head_net.features.23.bias", "pnp_net.features.1.weight")


def build_reference(cfg_rel: str):
    ref_gdrn = ref_shim.import_reference_gdrn()
    cfg = Config.fromfile(os.path.join(ref_shim.REFERENCE_ROOT, cfg_rel))
    cfg = postprocess_like_main_gdrn(cfg, device="cpu")
    cfg.MODEL.CDPN.BACKBONE.PRETRAINED = ""
    torch.manual_seed(0)
    model, optimizer = ref_gdrn.build_model_optimizer(copy.deepcopy(cfg))
    return model, optimizer, cfg


def grad_summary(named_grads: dict) -> dict:
    out = {}
    for k, g in named_grads.items():
        g = g.detach().double().flatten()
        out["gnorm/" + k] = np.array(float(g.norm()))
        if k in FULL_GRAD:
            out["gfull/" + k] = g.float().numpy()
        else:
            step = max(1, g.numel() // 257)
            out["gsample/" + k] = g[::step][:257].float().numpy()
    return out


def run_case(model, batch, train_bn: bool, do_loss: bool):
    model.train(train_bn)
    kw = synth.forward_kwargs(batch, train=do_loss)
    if do_loss:
        for p in model.parameters():
            p.grad = None
        out_dict, loss_dict = model(batch["roi_img"].clone(), **kw)
        total = sum(loss_dict.values())
        total.backward()
        grads = {k: p.grad for k, p in model.named_parameters()}
        return loss_dict, grads
    with torch.no_grad():
        return model(batch["roi_img"].clone(), **kw), None


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    report = []

    model, _opt, cfg = build_reference(REF_CFG)
    names = [k for k, _ in model.named_parameters()]
    assert len(names) == 148, len(names)
    raw = synth.seeded_state_dict(model.state_dict(), seed=0)
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "state_dict_manifest.npz"),
        names=np.array(list(raw.keys())),
        shapes=np.array([str(tuple(v.shape)) for v in raw.values()]),
        checksums=np.array([float(v.double().sum()) for v in raw.values()]),
        param_names=np.array(names),
    )
    sd = fixtures.calibrated_state_dict(0)  # needs the manifest written above
    model.load_state_dict(sd)

    # ---- case 1: eval forward, B=2 (BASELINE.json configs[0]) ------------------------------
    batch = synth.make_batch(2, seed=0)
    model.cfg.TEST.USE_PNP = True
    out, _ = run_case(model, batch, train_bn=False, do_loss=False)
    model.cfg.TEST.USE_PNP = False
    head_ref = torch.cat([out["mask"], out["coor_x"], out["coor_y"], out["coor_z"], out["region"]], dim=1)
    o = O.gdrn_forward(O.leaf_state_dict(sd, requires_grad=False), batch, train=False, do_loss=False)
    dev = {
        "head": float((o["head"] - head_ref).abs().max() / head_ref.abs().max()),
        "rot": float((o["rot"] - out["rot"]).abs().max()),
        "trans": float((o["trans"] - out["trans"]).abs().max()),
    }
    report.append(("eval_b2", dev))
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "eval_b2.npz"),
        head=head_ref.numpy(), rot=out["rot"].numpy(), trans=out["trans"].numpy(),
        region_argmax=head_ref[:, 4:].argmax(1).numpy().astype(np.uint8),
    )

    # ---- case 2: train-mode (batch-stat BN) forward + backward, B=4 ---------------------------
    batch = synth.make_batch(4, seed=1)
    model.load_state_dict(sd)
    loss_ref, grads_ref = run_case(model, batch, train_bn=True, do_loss=True)
    vis_ref = dict(ref_shim.event_storage().scalars)
    bn_after = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k}
    leaf = O.leaf_state_dict(sd)
    o = O.gdrn_forward(leaf, batch, train=True, do_loss=True, update_stats=True)
    sum(o["losses"].values()).backward()
    dev = {k: float(abs(o["losses"][k] - loss_ref[k]) / abs(loss_ref[k])) for k in loss_ref}
    gdev = 0.0
    for k in names:
        gr = grads_ref[k]
        gdev = max(gdev, float((leaf[k].grad - gr).norm() / (gr.norm() + 1e-30)))
    dev["grad_rel_l2_max"] = gdev
    dev["bn_running"] = max(float((leaf[k] - v).abs().max()) for k, v in bn_after.items())
    dev["vis/error_R"] = abs(o["vis"]["vis/error_R"] - vis_ref["vis/error_R"])
    report.append(("train_b4", dev))
    save = {("loss/" + k): np.array(float(v)) for k, v in loss_ref.items()}
    save.update(grad_summary(grads_ref))
    save.update({("vis/" + k.split("/")[1]): np.array(float(v)) for k, v in vis_ref.items()})
    for k in ("backbone.bn1.running_mean", "backbone.layer4.2.bn2.running_var", "rot_head_net.features.21.running_mean"):
        save["bn/" + k] = bn_after[k].numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "train_b4.npz"), **save)

    # ---- case 3: YCB-V style symmetric PM loss (PM_LOSS_SYM=True), B=4 -------------------------
    model_y, _o, cfg_y = build_reference(REF_CFG_YCBV)
    assert cfg_y.MODEL.CDPN.PNP_NET.PM_LOSS_SYM is True
    assert list(model_y.state_dict().keys()) == list(sd.keys())
    model_y.load_state_dict(sd)
    batch = synth.make_batch(4, seed=2, with_sym=True)
    loss_ref, grads_ref = run_case(model_y, batch, train_bn=True, do_loss=True)
    leaf = O.leaf_state_dict(sd)
    o = O.gdrn_forward(leaf, batch, train=True, do_loss=True, pm_sym=True)
    sum(o["losses"].values()).backward()
    dev = {k: float(abs(o["losses"][k] - loss_ref[k]) / abs(loss_ref[k])) for k in loss_ref}
    dev["grad_rel_l2_max"] = max(
        float((leaf[k].grad - grads_ref[k]).norm() / (grads_ref[k].norm() + 1e-30)) for k in names
    )
    report.append(("train_sym_b4", dev))
    save = {("loss/" + k): np.array(float(v)) for k, v in loss_ref.items()}
    save.update(grad_summary(grads_ref))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "train_sym_b4.npz"), **save)

    for name, dev in report:
        print(name, {k: f"{v:.3e}" for k, v in dev.items()})


if __name__ == "__main__":
    main()
